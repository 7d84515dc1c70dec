"""Drop-in for the reference module `optical_DP_channel/func_CMAbatch_DP_MQAM_shaping.py` (processing at line 15):
put this directory ahead of the reference's on sys.path and the unmodified Eval_run_*.py driver
imports this file instead.  Same positional signature and return values; runs on the CUDA path."""
import _path  # noqa: F401
from vae_equalizer_b200.processing import processing_cmabatch_dp as processing  # noqa: E402,F401
