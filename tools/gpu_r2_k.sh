#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/run_sweep.py tf32 1024 512 2 || { echo QUICK FAILED; exit 1; }
timeout 120 python tools/run_sweep.py tf32 2048 1024 2 || { echo QUICK FAILED; exit 1; }
cat > /tmp/t4.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import edrl_b200
for (N, d) in ((8192, 512), (8192, 768), (8192, 256), (8192, 1024), (16384, 1024), (4096, 512)):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
    y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
    for prec in ("tf32",):
        def step():
            x.grad = None; y.grad = None
            l = edrl_b200.MK_MMD(x, y, precision=prec); l.backward(); return l
        for _ in range(3): l = step()
        ts = []
        for _ in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"N={N} d={d} {prec}: {min(ts):.3f} ms  TF/s {12.0*N*N*d/min(ts)/1e9:.1f} loss {l.item():.8f}", flush=True)
PY
timeout 200 python /tmp/t4.py
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_mmd.py -q -m gpu -x 2>&1 | tail -3
