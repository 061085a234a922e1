"""A few launches of one select shape (ncu target): python tools/one_topk.py R W k sorted(0|1)"""
import sys, torch
sys.path.insert(0, ".")
import edrl_b200
R, W, k, srt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
x = torch.randn(R, W, device="cuda")
for _ in range(4):
    v, i = edrl_b200.topk_rows(x, k, sorted=srt)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    edrl_b200.topk_rows(x, k, sorted=srt)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"R={R} W={W} k={k} sorted={srt}: {ms:.4f} ms {(R*W*4+R*k*8)/ms/1e6:.0f} GB/s")
