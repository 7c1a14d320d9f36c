"""Key metrics of every launch in an .ncu-rep (read here, without a GPU): `python tools/ncu_summary.py rep [out.csv]`."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        f.write(raw)
for r in data:
    print("-" * 60)
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print("%-64s %s %s" % (h, r[i][:90], units[i]))
