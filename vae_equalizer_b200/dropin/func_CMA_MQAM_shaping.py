"""Drop-in for the reference module `AWGN_channel/func_CMA_MQAM_shaping.py` (processing at line 201, CMA :142, CPE :170,
SER_CMA :63, find_shift_symb :127): put this directory ahead of the reference's on sys.path and the unmodified
Eval_run_shaping_cma.py driver imports this file instead.  Same positional signatures and return values; runs on the CUDA path."""
import _path  # noqa: F401
from vae_equalizer_b200.awgn_cma import CMA, CPE, SER_CMA, find_shift_symb  # noqa: E402,F401
from vae_equalizer_b200.processing import processing_cma_awgn as processing  # noqa: E402,F401
