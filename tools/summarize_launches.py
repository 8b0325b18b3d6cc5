"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for ONE training step
(delimited by the fused Adam kernel)."""
import collections, csv, sys

def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))

def us(x):
    v = float(x["Metric Value"].replace(",", ""))
    return {"ns": v / 1e3, "nsecond": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}[x["Metric Unit"]]

def main(path):
    rows = load(path)
    idx = [i for i, r in enumerate(rows) if "adam" in r["Kernel Name"]]
    step = rows[idx[-3] + 1: idx[-2] + 1] if len(idx) >= 3 else rows
    tot = sum(us(r) for r in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        n = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
        agg[n][0] += 1
        agg[n][1] += us(r)
    print(f"one training step: {len(step)} launches, {tot / 1e3:.3f} ms of device time (cold-cache, serialised)")
    print(f"{'share':>7} {'ms':>9} {'count':>6}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / tot:6.1f}% {t / 1e3:9.3f} {c:6d}  {n}")

if __name__ == "__main__":
    main(sys.argv[1])
