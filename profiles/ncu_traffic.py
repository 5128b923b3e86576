"""DRAM traffic and pipe utilisation per kernel of an `ncu --set full` report, as JSON for bench.py (roofline.traffic,
roofline.step.traffic).  usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu_traffic.py raw.csv frames_per_launch out.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
frames = int(sys.argv[2])
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k, scale_units=True):
    v = r[ix[k]]
    if v == "": return None
    v = float(v)
    u = units[ix[k]]
    if scale_units:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v
roles = (("frame_sum", "prepass"), ("tile_prepare", "tile_records"), ("fbank512_v6", "main"), ("cmvn_utt_apply", "apply"))
out = {"workload": "bench.py default: 8192 ragged utterances (%d frames)" % frames, "kernels": {}}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    role = next((b for a, b in roles if a in name), None)
    if role is None or role in out["kernels"]:
        continue
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    k = {"kernel": name.split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
         "frames_per_launch": frames, "duration_ms_under_ncu": val(r, "gpu__time_duration.sum"),
         "registers_per_thread": val(r, "launch__registers_per_thread", False), "grid": val(r, "launch__grid_size", False),
         "block": val(r, "launch__block_size", False),
         "issue_active_pct": val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active", False) if "sm__inst_issued.avg.pct_of_peak_sustained_active" in ix else None,
         "pipe_fma_cycles_active_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
         "l1tex_data_pipe_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
         "pipe_lsu_pct": val(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", False),
         "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
         "tmem_ld_instructions": val(r, "smsp__sass_inst_executed_op_tmem_ldt.sum", False) if "smsp__sass_inst_executed_op_tmem_ldt.sum" in ix else None,
         "shared_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
         "warp_instructions": val(r, "smsp__inst_executed.sum", False)}
    out["kernels"][role] = k
alg = 960.0 * frames
tot = sum(k["traffic_bytes_per_launch"] for k in out["kernels"].values())
out["step_traffic_bytes"] = tot
out["algorithmic_bytes"] = alg
out["step_traffic_over_algorithmic"] = tot / alg
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps({k: (v["traffic_bytes_per_launch"], v["duration_ms_under_ncu"]) for k, v in out["kernels"].items()}), tot / alg)
