import sys, time, torch
sys.path.insert(0, ".")
import edrl_b200
from edrl_b200 import _lib
for (N, d) in ((64, 3072), (256, 512)):
    x = torch.randn(N, d, device="cuda"); y = torch.randn(N, d, device="cuda") * 1.25 + 0.1
    def step():
        a = x.detach().requires_grad_(True); b = y.detach().requires_grad_(True)
        edrl_b200.MK_MMD(a, b).backward()
    for _ in range(20): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): step()
    t1 = time.perf_counter()          # CPU enqueue time only
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"N={N} d={d}: cpu enqueue {1e6*(t1-t0)/200:.1f} us/step, total {1e6*(t2-t0)/200:.1f} us/step")
    # forward only / pieces
    lib = _lib.load()
    from edrl_b200.mmd import Workspace
    t0 = time.perf_counter()
    for _ in range(200): ws = Workspace(N, N, d, 0, x.device)
    print(f"   workspace alloc {1e6*(time.perf_counter()-t0)/200:.1f} us")
    loss = torch.empty((), device="cuda"); stats = torch.empty(8, device="cuda")
    st = _lib.stream_and_device(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        lib.edrl_mmd_forward(x.data_ptr(), y.data_ptr(), N, N, d, 2.0, 5, 0, 0, 1, loss.data_ptr(), stats.data_ptr(), None, ws.ptr, ws.nbytes, st)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"   C-ABI forward: cpu {1e6*(t1-t0)/200:.1f} us, total {1e6*(t2-t0)/200:.1f} us")
    g = torch.ones((), device="cuda"); dz = torch.empty(2*N, d, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        lib.edrl_mmd_backward(N, N, d, 2.0, 5, 0, stats.data_ptr(), g.data_ptr(), 0, 2*N, dz.data_ptr(), ws.ptr, ws.nbytes, st)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"   C-ABI backward: cpu {1e6*(t1-t0)/200:.1f} us, total {1e6*(t2-t0)/200:.1f} us")
