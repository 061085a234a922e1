"""MK_MMD forward+backward at the headline shape in one precision mode (for ncu captures of the sweep kernel)."""
import sys
import torch
sys.path.insert(0, ".")
import edrl_b200

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
g = torch.Generator(device="cuda").manual_seed(1013)
x = torch.randn(N, d, device="cuda", generator=g, requires_grad=True)
y = (torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1).requires_grad_(True)
for _ in range(iters):
    x.grad = None
    y.grad = None
    loss = edrl_b200.MK_MMD(x, y, precision=prec)
    loss.backward()
torch.cuda.synchronize()
print(prec, N, d, float(loss))
