"""Print selected metrics of every kernel in an .ncu-rep (reads `ncu -i rep --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "l1tex__data_pipe", "smsp__inst_executed.sum",
                        "sm__cycles_elapsed.max", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
                        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:70])
    for i, n in enumerate(h):
        if any(n == p or (p.endswith("*") and n.startswith(p[:-1])) or (p in n and len(p) > 12 and not p.endswith("*")) for p in pats):
            print(f"   {n} [{units[i]}] = {r[i]}")
