import sys, numpy as np, torch
sys.path.insert(0, ".")
import edrl_b200
from oracle import edrl_oracle as O
from oracle.gen_golden import mmd_inputs
for (seed, ns, nt, d, sh, sc) in ((104, 37, 53, 24, 0.2, 1.3), (105, 256, 256, 512, 0.1, 1.25), (1011, 2048, 2048, 512, 0.1, 1.25), (7, 300, 212, 700, 0.3, 2.0)):
    x, y = mmd_inputs(seed, ns, nt, d, sh, sc); x, y = x.numpy(), y.numpy()
    ref, _, dx, dy = O.mk_mmd_grad(x, y)
    gmax = max(np.abs(dx).max(), np.abs(dy).max())
    for prec in ("tf32", "tf32h", "f16s", "3xtf32"):
        xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
        yt = torch.tensor(y, dtype=torch.float32, device="cuda", requires_grad=True)
        l = edrl_b200.MK_MMD(xt, yt, precision=prec); l.backward()
        ex = np.abs(xt.grad.cpu().numpy() - dx).max() / gmax; ey = np.abs(yt.grad.cpu().numpy() - dy).max() / gmax
        rms = np.sqrt(np.mean((np.concatenate([xt.grad.cpu().numpy(), yt.grad.cpu().numpy()]) - np.concatenate([dx, dy])) ** 2)) / gmax
        print(f"{ns}x{nt}x{d} {prec:7s} loss rel err {abs(l.item()-ref)/ref:.2e}  grad max err/gmax {max(ex,ey):.2e}  rms/gmax {rms:.2e}")
# step timing of the modes at the headline shape
import time
N, d = 8192, 512
g = torch.Generator(device="cuda").manual_seed(1013)
x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for prec in ("tf32", "tf32h", "f16s"):
    def step():
        x.grad = None; y.grad = None
        edrl_b200.MK_MMD(x, y, precision=prec).backward()
    for _ in range(5): step()
    ts = []
    for _ in range(20):
        fl.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    print(f"{prec} step ms {ms:.3f} (min {min(ts):.3f})  algorithmic TF/s {12.0*N*N*d/ms/1e9:.1f}")
