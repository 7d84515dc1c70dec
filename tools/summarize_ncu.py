"""profiles/ helper: condense `ncu --page raw --csv` exports (one per .ncu-rep) and the launch list into the tracked summaries.
usage: python tools/summarize_ncu.py TAG raw1.csv[:label] [raw2.csv[:label] ...]   (writes profiles/TAG_ncu_full_summary.json)"""
import csv, json, sys
WANT = {'duration_us': 'gpu__time_duration.sum', 'dram_read_MB': 'dram__bytes_read.sum', 'dram_write_MB': 'dram__bytes_write.sum',
        'dram_pct_of_peak': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'fma_pipe_pct': 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'alu_pipe_pct': 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'xu_pipe_pct': 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'lsu_pipe_pct': 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'registers': 'launch__registers_per_thread', 'warp_inst': 'smsp__inst_executed.sum',
        'grid': 'launch__grid_size', 'block': 'launch__block_size', 'smem_dyn_KB': 'launch__shared_mem_per_block_dynamic',
        'stall_long_sb': 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'stall_not_selected': 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'stall_math_throttle': 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'stall_barrier': 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_wait': 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'stall_dispatch': 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'stall_short_sb': 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'}
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
if len(sys.argv) < 3 or sys.argv[1].startswith('-'):
    sys.exit(__doc__)
tag, out = sys.argv[1], {}
for arg in sys.argv[2:]:
    path, _, label = arg.partition(':')
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[idx['Kernel Name']].replace('void ', '').split('(')[0] + (f" [{label}]" if label else "")
        if name in out:
            continue
        rec = {}
        for k, m in WANT.items():
            if m in idx:
                v, u = float(r[idx[m]].replace(',', '')), units[idx[m]]
                if k in ('duration_us', 'dram_read_MB', 'dram_write_MB') and u in SCALE:
                    v *= SCALE[u]
                rec[k] = round(v, 3)
        out[name] = rec
json.dump(out, open(f'profiles/{tag}_ncu_full_summary.json', 'w'), indent=1)
for k, v in out.items():
    print(k, {a: v.get(a) for a in ('duration_us', 'dram_read_MB', 'dram_write_MB', 'dram_pct_of_peak', 'fma_pipe_pct', 'issue_active_pct', 'registers', 'grid', 'stall_long_sb', 'stall_barrier')})
