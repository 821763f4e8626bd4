"""Prints the handful of ncu metrics we track from a .ncu-rep (raw page) and the hottest SASS lines."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct", "smsp__cycles_active.avg", "sm__warps_active.avg.pct", "gpu__dram_throughput",
        "sm__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) or w in h for w in WANT) and "per_second" not in h or h in ("dram__bytes_read.sum.per_second",):
            if "pct_of_peak_sustained_elapsed" in h and not any(w in h for w in ("tensor", "lts__throughput", "dram__throughput", "sm__throughput")):
                continue
            print(f"   {h} [{u}] = {v}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))[2:]
    tot = sum(int(r[4] or 0) for r in srows)
    print("total samples", tot)
    for r in sorted(srows, key=lambda r: -int(r[4] or 0))[: int(sys.argv[2])]:
        print(f"   {r[4]:>8} {r[5]:>10}  {r[1].strip()[:110]}")
