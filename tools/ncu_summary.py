"""Summarise ncu reports (read here, no GPU): python tools/ncu_summary.py <rep> [<launches.csv>] > profiles/x.md"""
import csv, subprocess, sys, collections, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
# python tools/ncu_summary.py <rep> [<launches.csv>] [--traffic <json> <workload> <symbols>]: also records the kernel's
# DRAM bytes per launch in <json> (bench.py reports them as roofline.traffic for the same workload size)
traffic_args = None
if "--traffic" in sys.argv:
    k = sys.argv.index("--traffic")
    traffic_args = sys.argv[k + 1:k + 4]
    del sys.argv[k:k + 4]
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
print("## ncu --set full: %s\n" % rep.split('/')[-1])
for r in rows[2:]:
    print("kernel: `%s`\n" % r[H.index('Kernel Name')])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in H:
            print("| %s | %s | %s |" % (k, r[H.index(k)], U[H.index(k)]))
    tens = [h for h in H if 'tensor' in h or 'tmem' in h.lower() or 'utc' in h.lower()]
    for k in tens:
        if k not in KEYS and r[H.index(k)] not in ('', '0', 'n/a') and not any(x in k for x in ('.min', '.max', '.sum', 'ops_path', 'attribute')):
            print("| %s | %s | %s |" % (k, r[H.index(k)], U[H.index(k)]))
    print()
if traffic_args:
    import json, os
    def gb(v, unit):
        return float(v.replace(',', '')) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit]
    r = rows[2]
    rd = gb(r[H.index('dram__bytes_read.sum')], U[H.index('dram__bytes_read.sum')])
    wr = gb(r[H.index('dram__bytes_write.sum')], U[H.index('dram__bytes_write.sum')])
    path, wl, nsym = traffic_args
    db = json.load(open(path)) if os.path.exists(path) else {}
    db[wl] = {"kernel": r[H.index('Kernel Name')], "symbols_per_gpu": int(nsym), "dram_bytes_read": rd,
              "dram_bytes_write": wr, "traffic": rd + wr, "source": rep.split('/')[-1] + " (ncu --set full, one launch)"}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]; data = rows[hdr + 1:]
    ki, vi = H.index('Kernel Name'), H.index('Metric Value')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki].split('(')[0][:70], []).append(float(r[vi].replace(',', '')))
    print("## launch list (gpu__time_duration.sum, ns; cold-cache, serialised): %s\n" % sys.argv[2].split('/')[-1])
    print("| kernel | launches | mean ns | max ns | total ns |\n|---|---|---|---|---|")
    for k, v in agg.items():
        print("| `%s` | %d | %.0f | %.0f | %.0f |" % (k, len(v), sum(v) / len(v), max(v), sum(v)))
