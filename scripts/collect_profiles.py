"""Turn gpurun_out/<round>_* into the tracked summaries under profiles/ (run here, where ncu can read reports)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for name in ("bench.json", "bench_ref.json", "pytest_gpu.txt", "fused_nms_bench.txt"):
    src = os.path.join(G, f"{R}_{name}")
    if os.path.exists(src):
        open(os.path.join(P, f"{R}_{name}"), "w").write(open(src).read())

rows = list(csv.reader(open(os.path.join(G, f"{R}_launches.csv"), errors="ignore")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("pgmp::<unnamed>::", "").replace("void ", "")
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
ours = {k: v for k, v in agg.items() if not k.startswith("at::") and "nccl" not in k.lower()}
tot = sum(v[1] for v in ours.values())
with open(os.path.join(P, f"{R}_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 700 -- python bench.py --steps 2 --warmup 3\n")
    f.write("# (tc mode, 32 x 512px images, kNN-50).  Per-launch times under ncu are cold-cache and serialised:\n")
    f.write("# compare SHARES with bench.py's CUDA-event shares (profiles/%s_bench.json 'kernels'), not absolutes.\n" % R)
    f.write("# libpgmp kernels only (torch's input-generation / copy kernels excluded); unit = ns\n")
    for k, (c, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:44s} launches={c:4d} total_ns={t:14.0f} share={100 * t / tot:5.1f}%\n")

traffic = {}
for rep, note in ((f"{R}_edge_step_tc", "dominant kernel: one message-passing step over 929 518 edges (B=32)"),
                  (f"{R}_nms", "heatmap NMS + candidate extraction over 32 x 17 x 512 x 512 fp32"),
                  (f"{R}_group", "grouping tail (GAEC), one CTA per image, 32 images"),
                  (f"{R}_gather_features", "node-feature gather from NCHW maps (one 32-byte sector per element)"),
                  (f"{R}_nms_fused", "NMS with the scoremap assembly fused into its load stage (pgmp_gc_detect_fused), 32 x 17 x 512 x 512")):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), path,
                    os.path.join(P, rep + "_ncu.txt"), note], stdout=subprocess.DEVNULL)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, r = rr[0], rr[1], rr[2]
    def val(k):
        x = float(r[h.index(k)])
        return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u[h.index(k)]]
    kname = {"edge_step_tc": "edge_step_tc_kernel", "nms": "nms_candidates_kernel", "group": "group_kernel",
             "gather_features": "gather_features_kernel", "nms_fused": "nms_candidates_kernel (fused assembly)"}[rep[len(R) + 1:]]
    traffic[kname] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                      "edges_per_launch": 929518, "source": f"profiles/{rep}_ncu.txt (ncu --set full, one launch)"}
json.dump(traffic, open(os.path.join(P, f"{R}_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{R}_launches_summary.txt")).read())
print(json.dumps(traffic, indent=1))
