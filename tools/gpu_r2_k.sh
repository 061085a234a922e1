#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/t4.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import edrl_b200
for (N, d) in ((1024, 512), (8192, 512), (8192, 768), (8192, 256), (4096, 640)):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
    y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
    res = {}
    for prec in ("3xtf32", "tf32"):
        def step():
            x.grad = None; y.grad = None
            l = edrl_b200.MK_MMD(x, y, precision=prec); l.backward(); return l
        for _ in range(3): l = step()
        ts = []
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[prec] = (l.item(), x.grad.clone())
        print(f"N={N} d={d} {prec}: {min(ts):.3f} ms  loss {l.item():.8f}", flush=True)
    gm = res["tf32"][1].abs().max().item()
    print("   3x vs tf32: loss diff", abs(res["3xtf32"][0] - res["tf32"][0]), "grad diff / gmax", (res["3xtf32"][1] - res["tf32"][1]).abs().max().item() / gm, flush=True)
PY
timeout 120 python /tmp/t4.py
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_mmd.py -q -m gpu -x 2>&1 | tail -3
