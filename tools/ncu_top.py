"""Print the top stalled instructions of an .ncu-rep (source page) and a few whole-kernel numbers."""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); hdr, units, row = r[0], r[1], r[2]
want = ['gpu__time_duration.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
for h, u, v in zip(hdr, units, row):
    if h in want: print(f'{h:88s} {u:10s} {v}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']        # one section per captured launch
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
end = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
print(rows[starts[sec]][1])
hdr = rows[starts[sec] + 1]; data = [r for r in rows[starts[sec] + 2:end] if len(r) == len(hdr)]
iS, iI, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'warp instrs', sum(int(r[iI]) for r in data))
top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:ntop]
for k in sorted(top):
    r = data[k]
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{k:4d} {100*int(r[iS])/tot:5.1f}% {int(r[iI]):8d}  {r[iSrc].strip()[:58]:58s} {st}")
