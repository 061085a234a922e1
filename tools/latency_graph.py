"""Small shapes: MK_MMD forward+backward eager (host-bound?) against the same step replayed from a CUDA graph (GPU side)."""
import sys, torch
sys.path.insert(0, ".")
import edrl_b200
for (N, d) in ((64, 3072), (256, 512), (1024, 512), (2048, 512)):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
    y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
    def step():
        x.grad = None; y.grad = None
        l = edrl_b200.MK_MMD(x, y); l.backward(); return l
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): step()
    b.record(); torch.cuda.synchronize()
    eager = a.elapsed_time(b) / 50
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        x.grad = torch.zeros_like(x); y.grad = torch.zeros_like(y)
        def gstep():
            l = edrl_b200.MK_MMD(x, y); l.backward(); return l
        for _ in range(3): gstep()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        gstep()
    for _ in range(5): cg.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(50): cg.replay()
    b.record(); torch.cuda.synchronize()
    graph = a.elapsed_time(b) / 50
    n0 = edrl_b200.launch_count(); step(); n1 = edrl_b200.launch_count()
    print(f"N={N} d={d}: eager {eager*1e3:.1f} us, graph replay {graph*1e3:.1f} us, {n1-n0} launches", flush=True)
