"""Summarise an ncu report (read here, no GPU needed) into a small text file for profiles/.
usage: python tools/summarize_ncu.py gpurun_out/prof_conv.ncu-rep profiles/r1_conv_tc_ncu_summary.txt"""
import io
import subprocess
import sys

import pandas as pd

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
d = pd.read_csv(io.StringIO(raw)).iloc[1:].reset_index(drop=True)
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
cols = [c for w in want for c in d.columns if c == w]
with open(out, "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on   ({rep})\n")
    f.write("per-launch values; durations are cold-cache / serialised (compare shares, not absolutes)\n\n")
    for i in range(len(d)):
        f.write(f"--- launch {i}\n")
        for c in cols:
            f.write(f"{c:90s} {d[c][i]}\n")
        st = [(c.replace('smsp__average_warps_issue_stalled_', '').replace('_per_warp_active.pct', ''), float(d[c][i]))
              for c in d.columns if 'issue_stalled' in c and c.endswith('per_warp_active.pct')]
        st.sort(key=lambda r: -r[1])
        f.write("top warp stall reasons (% of warp-active): " + ", ".join(f"{n}={v:.1f}" for n, v in st[:6]) + "\n\n")
print("wrote", out)
