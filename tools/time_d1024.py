import sys, torch
sys.path.insert(0, ".")
import edrl_b200
for (N, d) in ((2048, 1024), (8192, 1024), (4096, 700)):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
    y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
    for prec in ("tf32", "f16s"):
        def step():
            x.grad = None; y.grad = None
            l = edrl_b200.MK_MMD(x, y, precision=prec); l.backward(); return l
        for _ in range(3): l = step()
        ts = []
        for _ in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = min(ts)
        print(f"N={N} d={d} {prec}: {ms:.3f} ms  algorithmic TF/s {12.0*N*N*d/ms/1e9:.1f}  loss {l.item():.6f} gsum {x.grad.abs().sum().item():.6e}")
import ctypes
lib = edrl_b200._lib.load()
plan = (ctypes.c_int * 10)()
lib.edrl_mmd_sweep_plan(8192, 8192, 1024, 0, 16384, 0, 0, plan)
print("plan N=8192 d=1024:", list(plan))
