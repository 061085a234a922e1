#!/bin/bash
mkdir -p gpurun_out
echo "== 3xtf32 fused: smoke"; timeout 120 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
import edrl_b200
from oracle import edrl_oracle as O
rng = np.random.default_rng(0)
for (ns, nt, d) in ((37, 53, 24), (300, 212, 96), (1024, 768, 512), (700, 650, 700)):
    x = rng.standard_normal((ns, d)); y = rng.standard_normal((nt, d)) * 1.3 + 0.2
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    yt = torch.tensor(y, dtype=torch.float32, device="cuda", requires_grad=True)
    l = edrl_b200.MK_MMD(xt, yt, precision="3xtf32"); l.backward()
    ref, _, dx, dy = O.mk_mmd_grad(xt.detach().cpu().double().numpy(), yt.detach().cpu().double().numpy())
    gm = max(np.abs(dx).max(), np.abs(dy).max())
    print(ns, nt, d, "loss rel", abs(l.item() - ref) / ref, "grad", max(np.abs(xt.grad.cpu().numpy() - dx).max(), np.abs(yt.grad.cpu().numpy() - dy).max()) / gm, flush=True)
PY
echo "== mmd tests"; timeout 900 python -m pytest tests/test_gpu_mmd.py -x -q -m gpu 2>&1 | tail -8
echo "== timing"; timeout 120 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import edrl_b200
g = torch.Generator(device="cuda").manual_seed(1013)
x = torch.randn(8192, 512, device="cuda", generator=g, requires_grad=True)
y = (torch.randn(8192, 512, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
for prec in ("tf32", "3xtf32"):
    for _ in range(3):
        x.grad = None; y.grad = None; edrl_b200.MK_MMD(x, y, precision=prec).backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        x.grad = None; y.grad = None; edrl_b200.MK_MMD(x, y, precision=prec).backward()
    b.record(); torch.cuda.synchronize()
    print(prec, a.elapsed_time(b) / 10, "ms/step")
PY
