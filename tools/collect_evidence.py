"""Copy the files tools/gpu_evidence.sh left in gpurun_out/ into profiles/ (round-2 names) and print the headline numbers.
Run here (no GPU): python tools/collect_evidence.py"""
import collections
import csv
import json
import shutil
import subprocess
import sys

G, P = "gpurun_out/", "profiles/"
subprocess.run([sys.executable, "tools/ncu_traffic.py", G + "r02_prof_step.ncu-rep", P + "r02_step_traffic.json",
                "python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-extras --no-graph"],
               check=True, stdout=subprocess.DEVNULL)
rows = list(csv.DictReader([l for l in open(G + "r02_render_launches.csv") if not l.startswith("==")]))
first = [i for i, r in enumerate(rows) if "raygen_equirect" in r["Kernel Name"]][-1]


def us(x):
    v = float(x["Metric Value"].replace(",", ""))
    return {"ns": v / 1e3, "nsecond": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}[x["Metric Unit"]]


step = rows[first:]
tot = sum(us(r) for r in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    n = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
    agg[n][0] += 1
    agg[n][1] += us(r)
out = [f"one PanoMipNeRF render of a 128x256 panorama (32768 rays, ONE forward, 64+64 samples, normals + 10x10 env + "
       f"shading): {len(step)} launches, {tot / 1e3:.3f} ms of device time (ncu, cold-cache, serialised)",
       f"{'share':>7} {'ms':>9} {'count':>6}  kernel"]
out += [f"{100 * t / tot:6.1f}% {t / 1e3:9.3f} {c:6d}  {n}" for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
open(P + "r02_render_launches.txt", "w").write("\n".join(out) + "\n")
for f in ("r02_train_step_launches.csv", "r02_train_step_launches.txt", "r02_render_launches.csv"):
    shutil.copy(G + f, P + f)
for src, dst in (("bench_train.json", "r02_bench_train_n1.json"), ("bench_render.json", "r02_bench_render_n1.json"),
                 ("bench_reference.json", "r02_bench_reference_arm.json")):
    open(P + dst, "w").write(open(G + src).read().strip().splitlines()[-1] + "\n")
shutil.copy(G + "micro.log", P + "r02_micro_memory_bound_kernels.jsonl")
shutil.copy(G + "fused_micro.log", P + "r02_fused_mlp_microbench.jsonl")
d = json.loads(open(P + "r02_bench_train_n1.json").read())
r = d["roofline"]
print(out[0])
print(open(P + "r02_train_step_launches.txt").readline().strip())
print("train", round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), d["clocks"])
print("fused", round(r["frac"], 3), {k: (round(v["kernel_ms_per_step"], 3), round(v["frac"], 3)) for k, v in r["programs"].items()})
print("wgrad", round(r["other_kernels"]["wgrad_batch_kernel"]["frac"], 3), "mlp_stage", round(r["mlp_stage"]["frac"], 3),
      "whole", round(r["whole_step"]["frac"], 3))
print({k: (round(d[k]["ms_per_step"], 2), round(d[k]["value"])) for k in ("c4", "c1", "render")})
print({k: v["frac_of_hbm_roofline"] for k, v in d["micro_kernels"].items()})
rr = json.loads(open(P + "r02_bench_render_n1.json").read())
print("render", round(rr["ms_per_step"], 2), round(rr["value"]), "e2e", round(rr["e2e"]["value"]), round(rr["roofline"]["frac"], 3),
      round(rr["roofline"]["whole_step"]["frac"], 3))
ref = json.loads(open(P + "r02_bench_reference_arm.json").read())
print("reference arm", round(ref["value"], 1), ref["unit"], ref["cpu_baseline"]["cores"], "cores")
