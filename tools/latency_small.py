import os, sys, time, torch
sys.path.insert(0, ".")
import edrl_b200
from edrl_b200 import _lib
from edrl_b200.mmd import Workspace
lib = _lib.load()
for (N, d) in ((64, 3072), (256, 512), (64, 512)):
    x = torch.randn(N, d, device="cuda"); y = torch.randn(N, d, device="cuda") * 1.25 + 0.1
    def step():
        a = x.detach().requires_grad_(True); b = y.detach().requires_grad_(True)
        edrl_b200.MK_MMD(a, b).backward()
    for _ in range(20): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"N={N} d={d}: step cpu enqueue {1e6*(t1-t0)/200:.1f} us, total {1e6*(t2-t0)/200:.1f} us")
    ws = Workspace(N, N, d, 0, x.device)
    loss = torch.empty((), device="cuda"); stats = torch.empty(8, device="cuda")
    st = _lib.stream_and_device(x); n = 2 * N
    g = torch.ones((), device="cuda"); dz = torch.empty(n, d, device="cuda"); u = torch.empty(8 * n, d, device="cuda")
    def timeit(name, fn):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200): fn()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"   {name}: cpu {1e6*(t1-t0)/200:.1f} us, total {1e6*(t2-t0)/200:.1f} us")
    timeit("forward      ", lambda: lib.edrl_mmd_forward(x.data_ptr(), y.data_ptr(), N, N, d, 2.0, 5, 0, 0, 1, loss.data_ptr(), stats.data_ptr(), None, ws.ptr, ws.nbytes, st))
    timeit("backward     ", lambda: lib.edrl_mmd_backward(N, N, d, 2.0, 5, 0, stats.data_ptr(), g.data_ptr(), 0, n, dz.data_ptr(), ws.ptr, ws.nbytes, st))
    timeit("forward_grad ", lambda: lib.edrl_mmd_forward_grad(x.data_ptr(), y.data_ptr(), N, N, d, 2.0, 5, 0, 0, n, 0, 0, 1, loss.data_ptr(), stats.data_ptr(), None, u.data_ptr(), ws.ptr, ws.nbytes, st))
    timeit("apply_grad   ", lambda: lib.edrl_mmd_apply_grad(N, N, d, 0, stats.data_ptr(), g.data_ptr(), u.data_ptr(), 0, n, 0, 0, dz.data_ptr(), ws.ptr, ws.nbytes, st))
