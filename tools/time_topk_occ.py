import sys, torch
sys.path.insert(0, ".")
import edrl_b200
peak = 6548.2
for (R, W, k) in ((1 << 18, 512, 64), (1 << 18, 800, 100), (1 << 18, 1024, 100), (1 << 17, 1600, 100), (1 << 17, 2048, 100), (1 << 15, 8192, 100), (1 << 16, 4096, 100)):
    x = torch.randn(R, W, device="cuda")
    for s in (False, True):
        for _ in range(3):
            edrl_b200.topk_rows(x, k, sorted=s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            edrl_b200.topk_rows(x, k, sorted=s)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        byts = R * W * 4 + R * k * 8
        print(f"R={R} W={W} k={k} sorted={s}: {ms:.3f} ms  {byts/ms/1e6/peak*100:.1f}% of HBM")
    sv, si = torch.sort(x[:2048], dim=1, descending=True, stable=True)
    v, i = edrl_b200.topk_rows(x[:2048], k, sorted=True)
    assert torch.equal(v, sv[:, :k]) and torch.equal(i.long(), si[:, :k])
    del x
