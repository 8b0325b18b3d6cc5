"""Summarise an `ncu --set full` capture of the tensor-core kernels of one training step into the JSON that
bench.py reads for `roofline.traffic` (profiles/rNN_step_traffic.json): DRAM bytes and tensor-pipe activity per launch,
grouped by program.  Run where `ncu` is installed (no GPU needed):
    python tools/ncu_traffic.py gpurun_out/r02_prof_step.ncu-rep profiles/r02_step_traffic.json "<command profiled>"
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def family(name: str) -> str:
    if "wgrad_batch_kernel" in name:
        return "wgrad_batch"
    if "mlp_fused_kernel" in name:
        prog = name.split("mlp_fused_kernel<")[1].split(",")[0].strip().lstrip("(int)")
        return {"0": "mlp_fused", "1": "mlp_fused", "2": "mlp_fused_bwd", "3": "mlp_fused_jadj"}.get(prog, "mlp_fused")
    return name


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}

    def val(r, key):
        i = col[key]
        x = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
        return x * UNIT.get(units[i], 1.0)

    tkey = next(k for k in head if k.startswith("sm__pipe_tensor") and "cycles_active" in k and "pct" in k)
    fam = {}
    for r in body:
        f = fam.setdefault(family(r[col["Kernel Name"]]), {"launches_per_step": 0, "ncu_ms_per_step": 0.0,
                                                          "dram_read_bytes_per_step": 0.0,
                                                          "dram_write_bytes_per_step": 0.0,
                                                          "tensor_pipe_active_pct_per_launch": []})
        f["launches_per_step"] += 1
        f["ncu_ms_per_step"] += val(r, "gpu__time_duration.sum")
        f["dram_read_bytes_per_step"] += val(r, "dram__bytes_read.sum")
        f["dram_write_bytes_per_step"] += val(r, "dram__bytes_write.sum")
        f["tensor_pipe_active_pct_per_launch"].append(float(r[col[tkey]]))
    total = 0.0
    for f in fam.values():
        f["dram_bytes_per_step"] = f["dram_read_bytes_per_step"] + f["dram_write_bytes_per_step"]
        f["dram_bytes_per_launch"] = f["dram_bytes_per_step"] / f["launches_per_step"]
        total += f["dram_bytes_per_step"]
    fam["_total_dram_bytes_per_step"] = total
    fam["_tensor_metric"] = tkey
    fam["_source"] = f"ncu --set full --clock-control none on `{cmd}` (B200): the tensor-core kernels of one training step; tools/gpu_evidence.sh + tools/ncu_traffic.py"
    json.dump(fam, open(out, "w"), indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: v[kk] for kk in ("launches_per_step", "ncu_ms_per_step", "dram_bytes_per_step")}) for k, v in fam.items()}, indent=1))


if __name__ == "__main__":
    main()
