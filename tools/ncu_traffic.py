"""Builds profiles/ncu_traffic.json (DRAM bytes per launch of each kernel family of the C2 step) from the raw-page CSV of an
`ncu --set full` capture (tools/gpu_profile.sh writes gpurun_out/<tag>/ncu_full_top_kernels.csv).
usage: python tools/ncu_traffic.py <ncu_full_top_kernels.csv> <source tag>"""
import csv, json, sys, collections
from pathlib import Path

LABELS = [  # kernel-name fragment -> ops label of the family it belongs to (a family = one C-ABI call)
    ("knn2_", "lpd_knn_tc[C=64,k=20]"), ("knn_split", "lpd_knn_tc[C=64,k=20]"), ("knn_tc_kernel", "lpd_knn_tc[C=64,k=20]"),
    ("knn_refine", "lpd_knn_tc[C=64,k=20]"), ("edgeconv_dg_tc", "lpd_edgeconv_dg_tf32[128x128]"),
    ("edgeconv_dg20_h", "lpd_edgeconv_dg_f16[128x128]"),
    ("knn_grid_lockstep", "lpd_knn_xyz[k=20]"), ("knn_xyz_rescue", "lpd_knn_xyz[k=20]"), ("knn_grid_search", "lpd_knn_xyz[k=20]"),
    ("edge_gather_ext", "lpd_edge_gather_ext[C=256]"), ("edge_gather_max_h", "lpd_edge_gather_ext[C=256]"),
    ("pointwise_mlp2", "lpd_pointwise_mlp2"),
    ("gemm_tf32_kernel<64, 8, 0, __half, 0, 1>", "lpd_gemm_softmax64[262144x64x1024]"),
    ("gemm_tf32_kernel<64, 8, 1, __half, 0, 0>", "lpd_gemm_f16_tn[1024x64x4096x64]"),
    ("gemm_tf32_kernel<256, 4, 1, float, 0, 0>", "lpd_gemm_tf32_tn[64x256x1536x128]"),
    ("gemm_h2_kernel<float>", "lpd_gemm_tf32[262144x256x64]"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
per_kernel = collections.defaultdict(list)
for r in rows[2:]:
    if len(r) <= max(rd, wr) or not r[rd]:
        continue
    per_kernel[r[kn]].append(float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]])
fam = collections.defaultdict(float)
detail = collections.defaultdict(dict)
for name, vals in per_kernel.items():
    for frag, label in LABELS:
        if frag in name:
            mean = sum(vals) / len(vals)
            fam[label] += mean
            detail[label][name.split("(")[0][-60:]] = round(mean)
            break
out = {k: {"dram_bytes_per_launch": round(v), "kernels": detail[k], "source": f"profiles/{sys.argv[2]}_ncu_full_top_kernels.csv"}
       for k, v in fam.items()}
Path("profiles/ncu_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
