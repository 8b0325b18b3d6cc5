import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from util import O
from panonerf_b200 import ops
DEV="cuda"
gen = torch.Generator().manual_seed(0)
mean = (torch.rand(4096, 3, generator=gen) * 10 - 5)
cov = torch.rand(4096, 3, generator=gen) * torch.tensor([1e-6, 1e-4, 1e-2])
ref = O.ipe(mean, cov, 0, 16)
md, cd = mean.to(DEV), cov.to(DEV)
worst = 0
for it in range(200):
    out = torch.empty(4096, 96, device=DEV)
    ops.ipe_into(md, cd, 0, 16, out)
    d = (out.cpu() - ref).abs()
    e = float(d.max())
    if e > 2e-6:
        i = int(d.argmax()); m, j = divmod(i, 96)
        print("iter", it, "err", e, "sample", m, "feature", j, "out", float(out[m, j]), "ref", float(ref[m, j]), "mean", mean[m].tolist(), "cov", cov[m].tolist())
        worst += 1
        if worst > 5: break
print("done, bad iterations:", worst)
