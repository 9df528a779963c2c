"""Print the key metrics of an .ncu-rep (first kernel) — used to produce the summaries under profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum ', 'imma_cycles_active_realtime', 'lts__t_bytes.sum ', 'lts__throughput.avg.pct',
        'sm__cycles_elapsed.avg ', 'sm__cycles_active.avg ', 'smsp__cycles_active.avg ', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled', 'smsp__warp_issue_stalled', 'sm__clock', 'gpc__cycles_elapsed.max']
for h, u, v in zip(hdr, units, vals):
    if any(k in h + ' ' for k in keys) and 'per_second' not in h and '.min' not in h and '.max.' not in h:
        print(f'{h} [{u}] = {v}')
