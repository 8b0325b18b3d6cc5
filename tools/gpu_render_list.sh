mkdir -p gpurun_out
timeout 300 python bench.py --workload render --render-hw 128 256 --steps 1 --warmup 1 > gpurun_out/plain_r.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_render.csv \
    python bench.py --workload render --render-hw 128 256 --steps 1 --warmup 1 > gpurun_out/ncu_r.log 2>&1
echo "list rc $?"
python - <<'PY'
import collections, csv
lines=[l for l in open('gpurun_out/launches_render.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def us(x):
    v=float(x["Metric Value"].replace(",","")); return {"ns":v/1e3,"nsecond":v/1e3,"us":v,"usecond":v,"ms":v*1e3}[x["Metric Unit"]]
# last third of the launches = one render (warmup + timed + e2e)
n=len(rows)//3; step=rows[-n:]
tot=sum(us(r) for r in step); agg=collections.defaultdict(lambda:[0,0.0])
for r in step:
    k=r["Kernel Name"].split("(")[0].replace("void ","")[:70]; agg[k][0]+=1; agg[k][1]+=us(r)
print(f"one render of 32768 rays: {len(step)} launches, {tot/1e3:.3f} ms")
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print(f"{100*t/tot:6.1f}% {t/1e3:9.3f} {c:5d}  {k}")
PY
