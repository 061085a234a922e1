import sys, torch
sys.path.insert(0, ".")
import edrl_b200
for (R, W, k) in ((3, 100, 100), (8, 800, 100)):
    g = torch.Generator().manual_seed(R * 7 + W)
    x = torch.randn(R, W, generator=g)
    tv, ti = torch.topk(x, k, dim=1)
    for srt in (False, True):
        v, i = edrl_b200.topk_rows(x.cuda(), k, sorted=srt)
        v, i = v.cpu(), i.cpu().long()
        if not srt:
            o = torch.argsort(v, dim=1, descending=True, stable=True)
            v, i = torch.gather(v, 1, o), torch.gather(i, 1, o)
        bad = (v != tv) | (i != ti)
        print(R, W, k, "sorted" if srt else "unsorted", "mismatches:", int(bad.sum()))
        if bad.any():
            rr, cc = torch.nonzero(bad)[0].tolist()
            print("  first at row", rr, "col", cc, "got", v[rr, cc].item(), i[rr, cc].item(), "want", tv[rr, cc].item(), ti[rr, cc].item())
            print("  got vals", v[rr, max(0,cc-2):cc+4].tolist(), "idx", i[rr, max(0,cc-2):cc+4].tolist())
            print("  want    ", tv[rr, max(0,cc-2):cc+4].tolist(), "idx", ti[rr, max(0,cc-2):cc+4].tolist())
