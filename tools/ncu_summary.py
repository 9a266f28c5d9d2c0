#!/usr/bin/env python
"""Summarise an .ncu-rep (one profiled kernel) into a small text file for profiles/ and print DRAM traffic.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_ncu.txt"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
seen, lines, d = set(), [], {}
for h0, u, v in zip(hdr, units, vals):
    # the full set prefixes some columns with their section ("TPC.TriageCompute.sm__pipe_tensor_..."): match on the metric name
    h = h0.split("TriageCompute.")[-1]
    if h in keep and h not in seen:
        seen.add(h); lines.append(f"{h:88s} {u:16s} {v}"); d[h] = (u, v)
if "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg" in d and "sm__cycles_elapsed.avg" in d:
    # 4 tensor sub-pipes per SM: busy sub-pipe cycles / (4 x elapsed cycles) = share of the elapsed time the tensor pipe is issuing
    try:
        frac = float(d["sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"][1]) / (4.0 * float(d["sm__cycles_elapsed.avg"][1]))
        lines.append(f"{'tensor pipe active = hmma sub-pipe cycles / (4 x elapsed cycles)':88s} {'%':16s} {100 * frac:.2f}")
    except ValueError:
        pass
def bytes_of(k):
    u, v = d[k]; v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
traffic = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
lines.append(f"{'traffic = dram read + write per launch':88s} {'byte':16s} {traffic:.0f}")
open(out, "w").write(f"# ncu --set full --clock-control none, one launch of the kernel; source: {rep}\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
