"""Instruction mix and stall hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: python tools/ncu_sass_hist.py source.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, stall, tot_i, tot_s = collections.Counter(), collections.Counter(), 0, 0
recs = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG")) else ".".join(op.split(".")[:1])
    ie, ss = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    ops[op] += ie; stall[op] += ss; tot_i += ie; tot_s += ss
    recs.append((ss, ie, n, sass, r))
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("opcode            instr      share   samples  share")
for op, c in ops.most_common(28):
    print(f"{op:14s} {c:10d}  {c / tot_i:6.3f}  {stall[op]:8d}  {stall[op] / max(tot_s, 1):6.3f}")
print("\nhottest instructions by stall samples (index, samples, executed, sass, top stall reasons)")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for ss, ie, n, sass, r in sorted(recs, reverse=True)[:topn]:
    rs = sorted(((int(r[ix[k]]), k[6:]) for k in reasons), reverse=True)[:3]
    print(f"{n:6d} {ss:7d} {ie:9d}  {sass[:70]:70s} " + " ".join(f"{k}={v}" for v, k in rs if v))
