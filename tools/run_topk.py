import sys, torch
sys.path.insert(0, ".")
import edrl_b200
x = torch.randn(1 << 18, 800, device="cuda")
for s in (False, True):
    for _ in range(3):
        edrl_b200.topk_rows(x, 100, sorted=s)
torch.cuda.synchronize()
print("ok")
