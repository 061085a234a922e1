import sys, torch
sys.path.insert(0, ".")
import edrl_b200
B, T, Fd, S, C = 64, 216, 256, 800, 2
z = torch.randn(B, T, Fd, device="cuda"); prox = torch.randn(C, 2 * Fd, device="cuda") * 0.1
eps = torch.randn(C, S, Fd, device="cuda"); y = torch.randint(0, 2, (B,), device="cuda")
for _ in range(3):
    zz = z.detach().requires_grad_(True); pp = prox.detach().requires_grad_(True)
    att, _ = edrl_b200.essence_scores(zz, pp[:, :Fd], torch.nn.functional.softplus(pp[:, Fd:]), eps)
    loss, _, _ = edrl_b200.essence_select_loss(att, y, 100, sorted=False)
    loss.backward()
torch.cuda.synchronize()
print("ok")
