"""profiles/traffic.json: DRAM bytes per C-ABI call (ncu dram__bytes_read.sum + dram__bytes_write.sum) from the per-kernel
summary of one cfg2 forward (tools/step_metrics_summary.py); read by bench.py for the `traffic` field of its rooflines."""
import json, sys
src = sys.argv[1] if len(sys.argv) > 1 else "profiles/step_metrics_r1_v8.json"
k = json.load(open(src))["kernels"]

def fam(pred):
    n = sum(v["launches"] for name, v in k.items() if pred(name))
    b = sum(v["dram_bytes_per_launch"] * v["launches"] for name, v in k.items() if pred(name))
    return n, b

n_lin, b_lin = fam(lambda s: "linear_tc_kernel" in s or "linear_simt_kernel" in s)
n_fl, b_fl = fam(lambda s: "flash_attn_bf16_kernel" in s); _, b_mg = fam(lambda s: "flash_merge" in s)
n_g, b_g = fam(lambda s: "gat_edge_tc_kernel" in s); _, b_gf = fam(lambda s: "gat_finalize" in s)
n_p, b_p = fam(lambda s: "pointnet_tc_kernel" in s)
n_na, b_na = fam(lambda s: "node_attn_scene_kernel" in s)
t = {"_source": f"{src}: ncu dram__bytes_read.sum + dram__bytes_write.sum of one cfg2 forward (tools/prof_step.py), per C-ABI call "
                "(helper kernels of a call included; linear = mean over the 59 projections of a forward)",
     "vlsat_linear_fwd": round(b_lin / n_lin), "vlsat_flash_attn_bf16x3_fwd": round((b_fl + b_mg) / n_fl),
     "vlsat_gat_edge_tc_fwd": round((b_g + b_gf) / n_g), "vlsat_pointnet_tc_fwd": round(b_p / n_p),
     "vlsat_node_attn_scene_fwd": round(b_na / n_na)}
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
print(t)
