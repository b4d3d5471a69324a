"""Summarise an ncu report: `python profiles/ncu_summary.py X.ncu-rep > summary.csv` (one row per profiled launch)."""
import csv, io, subprocess, sys
M = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
     "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [m for m in M if m in ix]
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name", "Grid Size"] + cols)
w.writerow(["", ""] + [units[ix[c]] for c in cols])
for r in rows[2:]:
    w.writerow([r[ix["Kernel Name"]][:60], r[ix["Grid Size"]]] + [r[ix[c]] for c in cols])
