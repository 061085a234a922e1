"""Small shapes of every hot kernel, as a quick smoke run (compute-sanitizer is closed on this GPU pool)."""
import sys, torch
sys.path.insert(0, ".")
import edrl_b200
torch.manual_seed(0)
for prec in ("tf32", "tf32h", "f16s", "3xtf32"):
    for (ns, nt, d) in ((300, 212, 96), (1024, 768, 512), (130, 70, 1030)):
        x = torch.randn(ns, d, device="cuda", requires_grad=True)
        y = (torch.randn(nt, d, device="cuda") * 1.2 + 0.1).requires_grad_(True)
        l = edrl_b200.MK_MMD(x, y, precision=prec)
        l.backward()
        with torch.no_grad():
            edrl_b200.MK_MMD(x, y, precision=prec)
for (R, W, k) in ((37, 800, 100), (9, 1600, 100), (5, 216, 32), (3, 100, 7), (4, 2048, 128), (3, 5000, 100), (6, 333, 20)):
    x = torch.randn(R, W, device="cuda")
    x[0, : W // 2] = 0.5
    for s in (True, False):
        edrl_b200.topk_rows(x, k, sorted=s)
B, T, Fd, S, C = 8, 216, 256, 800, 2
z = torch.randn(B, T, Fd, device="cuda", requires_grad=True)
prox = (torch.randn(C, 2 * Fd, device="cuda") * 0.1).requires_grad_(True)
eps = torch.randn(C, S, Fd, device="cuda")
yl = torch.randint(0, 2, (B,), device="cuda")
edrl_b200.essence_train_loss(z, prox, eps, yl, 100).backward()
feat = torch.randn(16, 216, 768, device="cuda", requires_grad=True)
sc = torch.randn(16, 216, device="cuda")
out = edrl_b200.select_gather(feat, sc, 32)
out[0].sum().backward()
torch.cuda.synchronize()
print("small shapes done")
