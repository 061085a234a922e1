#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/t3.py <<'PY'
import sys, torch, ctypes
sys.path.insert(0, ".")
import edrl_b200
for (N, d) in ((8192, 1024), (16384, 1024), (32768, 1024), (8192, 2048), (8192, 1536), (4096, 1024)):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
    y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
    def step():
        x.grad = None; y.grad = None
        l = edrl_b200.MK_MMD(x, y); l.backward(); return l
    for _ in range(3): l = step()
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = min(ts)
    print(f"N={N} d={d}: {ms:.3f} ms  TF/s {12.0*N*N*d/ms/1e9:.1f}  loss {l.item():.7f} gsum {x.grad.abs().sum().item():.6e}", flush=True)
PY
for h in 0 1; do echo "== EDRL_MMD_HYBRID=$h"; EDRL_MMD_HYBRID=$h timeout 300 python /tmp/t3.py; done
echo "== tests default"; timeout 600 python -m pytest tests/test_gpu_mmd.py -q -m gpu -x 2>&1 | tail -3
echo "== tests forced hybrid"; EDRL_MMD_HYBRID=2 timeout 600 python -m pytest tests/test_gpu_mmd.py -q -m gpu -x 2>&1 | tail -3
