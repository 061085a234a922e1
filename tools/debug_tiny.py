import sys, numpy as np, torch
sys.path.insert(0, ".")
import edrl_b200
from oracle import edrl_oracle as O
for (ns, nt, d) in ((1, 1, 1), (1, 1, 8), (2, 2, 1), (3, 2, 2)):
    rng = np.random.default_rng(ns * 7919 + nt * 31 + d)
    x = rng.standard_normal((ns, d)) * 0.7 - 0.1
    y = rng.standard_normal((nt, d)) * 1.1 + 0.3
    for prec in ("3xtf32", "tf32"):
        xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
        yt = torch.tensor(y, dtype=torch.float32, device="cuda", requires_grad=True)
        l = edrl_b200.MK_MMD(xt, yt, precision=prec); l.backward()
        ref, m, dx, dy = O.mk_mmd_grad(xt.detach().cpu().numpy().astype(np.float64), yt.detach().cpu().numpy().astype(np.float64))
        l2, st = edrl_b200.mk_mmd_with_stats(xt.detach(), yt.detach(), precision=prec)
        print(ns, nt, d, prec, "loss", l.item(), ref, "gx", xt.grad.flatten()[:3].tolist(), dx.flatten()[:3].tolist(), "stats", st[:5].tolist())
