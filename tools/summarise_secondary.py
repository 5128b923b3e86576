"""Key figures of the raw ncu pages written by tools/profile_secondary.sh -> profiles/<round>_secondary_kernels_ncu.txt"""
import csv, glob, os, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
want = [("duration", "gpu__time_duration.sum"), ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
        ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"), ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("l1tex_data_pipe_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("warp_instr", "smsp__inst_executed.sum")]
out = ["ncu --set full --clock-control none, one launch per kernel (tools/profile_secondary.sh %s); units as reported by ncu" % R]
for f in sorted(glob.glob("gpurun_out/%s_sec_*.csv" % R)):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    h, u, r = rows[0], rows[1], rows[2]
    ix = {n: i for i, n in enumerate(h)}
    name = r[ix["Kernel Name"]].split("(")[0]
    parts = []
    for label, key in want:
        if key in ix and r[ix[key]] != "":
            parts.append("%s=%s%s" % (label, r[ix[key]], (" " + u[ix[key]]) if u[ix[key]] else ""))
    out.append("%s   [%s]" % (name, os.path.basename(f)[len(R) + 5:-4]))
    out.append("   " + ", ".join(parts))
open("profiles/%s_secondary_kernels_ncu.txt" % R, "w").write("\n".join(out) + "\n")
print("\n".join(out))
