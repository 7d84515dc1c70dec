"""Instruction mix of the hot kernels of libvaeq.so from `cuobjdump -sass` (no GPU needed): opcode histogram per kernel and the
mnemonics that prove which hardware paths a kernel uses (FFMA2 = packed fp32, LDGSTS = cp.async, UBLKCP = cp.async.bulk (TMA),
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, SYNCS = mbarrier).  usage: python tools/sass_mix.py [regex of kernel names] > profiles/..."""
import collections, re, subprocess, sys, os
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vae_equalizer_b200", "libvaeq.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_dp_|k_cma|k_er_|k_ser|k_awgn|k_gen|k_cpe|k_soft")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, mix = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = dem.split("(")[0].replace("void ", "") if pat.search(dem) else None
        if name:
            mix[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        mix[name][m.group(1)] += 1
KEY = ["FFMA2", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "UTCHMMA", "LDTM", "SYNCS", "BAR", "STL", "LDL", "SHFL"]
print("# static SASS instruction counts per kernel (cuobjdump -sass vae_equalizer_b200/libvaeq.so); STL / LDL = local-memory spills")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in KEY))
for n, c in mix.items():
    print(f"{n[:58]:58s} {sum(c.values()):6d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEY))
