"""GPU parity tests, Part A (MK_MMD): the sm_100a kernels, called through the C-ABI / the host
package, against the numpy oracle and the committed reference-generated golden vectors.

Tolerances (written here as the contract):
  3xTF32 mode : loss rtol 1e-4 (+ atol 1e-6), gradients 1e-4 * |grad|_inf absolute
  TF32 mode   : loss rtol 1e-3 (+ atol 1e-6), gradients 2e-3 * |grad|_inf absolute
  TF32H mode  : the same bar as TF32 (TF32 Gram; G.Z operands as scaled binary16 with the same 11-bit significands)
  F16S mode   : the same bar as TF32 (the Gram too reads a scaled binary16 copy of the TF32-rounded operand)
against the fp64 reference/oracle value.
"""
import os

import numpy as np
import pytest
import torch

from gpu_util import RawMMD, dev, have_gpu, tf32_round
from oracle import edrl_oracle as O
from oracle.gen_golden import MMD_CASES, MMD_VARIANTS, mmd_inputs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]

MODES = [("3xtf32", 1, 1e-4, 1e-4), ("tf32", 0, 1e-3, 2e-3), ("tf32h", 2, 1e-3, 2e-3), ("f16s", 4, 1e-3, 2e-3)]   # name, flag, loss rtol, grad atol / |g|_inf


@pytest.fixture(scope="module")
def raw():
    return RawMMD()


@pytest.fixture(scope="module")
def gm(golden_dir):
    return np.load(os.path.join(golden_dir, "mmd_reference.npz"))


def _case(case):
    x, y = mmd_inputs(*case)
    return x.numpy(), y.numpy()


# ---------------------------------------------------------------- plumbing: TMA -> tcgen05 -> TMEM
@pytest.mark.parametrize("ns,nt,d", [(4, 4, 8), (37, 53, 24), (128, 128, 32), (130, 200, 100), (256, 256, 512)])
@pytest.mark.parametrize("flag", [0, 1])
def test_centred_gram_tiles(raw, ns, nt, d, flag):
    rng = np.random.default_rng(ns * 1000 + d)
    x = rng.standard_normal((ns, d)).astype(np.float32)
    y = (rng.standard_normal((nt, d)) * 1.3 + 0.2).astype(np.float32)
    g = raw.kernel_matrix(dev(x), dev(y), flags=flag | 0x100).cpu().numpy().astype(np.float64)
    z = np.concatenate([x, y]).astype(np.float64)
    zc = (z - z.mean(0)).astype(np.float32)
    if flag == 0:
        zr = tf32_round(zc).astype(np.float64)
        ref = zr @ zr.T
        tol = 2e-5
    else:
        ref = zc.astype(np.float64) @ zc.astype(np.float64).T
        tol = 2e-5
    scale = np.abs(ref).max()
    # the device centres with an fp64 column mean rounded to fp32; allow that rounding too
    assert np.abs(g - ref).max() <= tol * scale + 1e-6 * scale, np.abs(g - ref).max() / scale
    # symmetric up to the summation order of the hi/lo cross terms inside a diagonal tile
    assert np.abs(g - g.T).max() <= (0.0 if flag == 0 else 1e-6) * scale


@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
def test_gaussian_kernel_matrix(raw, mode):
    import edrl_b200
    x, y = _case(MMD_CASES[3])
    k = edrl_b200.gaussian_kernel(dev(x), dev(y), precision=mode[0]).cpu().numpy()
    ref = O.gaussian_kernel(x, y)
    assert k.shape == ref.shape
    np.testing.assert_allclose(k, ref, rtol=mode[2] * 10, atol=mode[2])
    # hand-checkable case of SURVEY.md section 4
    xh = np.array([[0.0], [1.0]]); yh = np.array([[2.0], [3.0]])
    kh = edrl_b200.gaussian_kernel(dev(xh), dev(yh), precision=mode[0]).cpu().numpy()
    np.testing.assert_allclose(kh[0], [5.0, 3.37927553, 1.68977177, 0.84013917], rtol=2e-3)


# ---------------------------------------------------------------- known-answer table (reference-generated)
@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
@pytest.mark.parametrize("case", MMD_CASES, ids=lambda c: f"seed{c[0]}")
def test_mk_mmd_kat_forward_backward(case, mode, gm):
    import edrl_b200
    name, flag, ltol, gtol = mode
    seed = case[0]
    x, y = _case(case)
    xt = dev(x).requires_grad_(True)
    yt = dev(y).requires_grad_(True)
    loss = edrl_b200.MK_MMD(xt, yt, precision=name)
    assert loss.dim() == 0 and loss.device.type == "cuda" and loss.dtype == torch.float32
    loss.backward()
    ref = float(gm[f"s{seed}_loss64"])
    assert np.isclose(loss.item(), ref, rtol=ltol, atol=1e-6), (loss.item(), ref)
    _, _, dx, dy = O.mk_mmd_grad(x, y)
    gmax = float(gm[f"s{seed}_maxabs_g64"])
    gx = xt.grad.cpu().numpy().astype(np.float64)
    gy = yt.grad.cpu().numpy().astype(np.float64)
    assert np.abs(gx - dx).max() <= gtol * gmax, np.abs(gx - dx).max() / gmax
    assert np.abs(gy - dy).max() <= gtol * gmax, np.abs(gy - dy).max() / gmax
    # the committed reference gradients themselves (strided subsample)
    rs, rt, cs = gm[f"s{seed}_sub_strides"]
    assert np.abs(gx[::rs, ::cs] - gm[f"s{seed}_gx64_sub"]).max() <= gtol * gmax
    assert np.abs(gy[::rt, ::cs] - gm[f"s{seed}_gy64_sub"]).max() <= gtol * gmax
    assert np.isclose(np.abs(gx).sum(), float(gm[f"s{seed}_sumabs_gx64"]), rtol=max(gtol, 1e-3))


@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
@pytest.mark.parametrize("mul,num", MMD_VARIANTS)
def test_mk_mmd_kernel_variants(mul, num, mode, gm):
    import edrl_b200
    name, flag, ltol, gtol = mode
    x, y = _case(MMD_CASES[3])
    xt = dev(x).requires_grad_(True)
    yt = dev(y).requires_grad_(True)
    loss = edrl_b200.MK_MMD(xt, yt, kernel_mul=mul, kernel_num=num, precision=name)
    (2.5 * loss).backward()
    key = f"var_m{mul}_k{num}"
    assert np.isclose(loss.item(), float(gm[key + "_loss"]), rtol=ltol, atol=1e-6)
    gmax = max(np.abs(gm[key + "_gx"]).max(), np.abs(gm[key + "_gy"]).max()) * 2.5
    assert np.abs(xt.grad.cpu().numpy() - 2.5 * gm[key + "_gx"]).max() <= gtol * gmax
    assert np.abs(yt.grad.cpu().numpy() - 2.5 * gm[key + "_gy"]).max() <= gtol * gmax


@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
def test_mk_mmd_edge_cases(mode, gm):
    import edrl_b200
    name, flag, ltol, gtol = mode
    # hand case
    xh = dev(np.array([[0.0], [1.0]])); yh = dev(np.array([[2.0], [3.0]]))
    assert np.isclose(edrl_b200.MK_MMD(xh, yh, precision=name).item(), 4.579796409474792, rtol=ltol)
    # identical inputs: loss 0, zero gradient (up to fp32 cancellation)
    xi, _ = mmd_inputs(7, 6, 6, 5, 0, 1)
    a = dev(xi.numpy()).requires_grad_(True)
    b = dev(xi.numpy()).requires_grad_(True)
    l = edrl_b200.MK_MMD(a, b, precision=name)
    l.backward()
    assert abs(l.item()) <= 1e-6
    assert a.grad.abs().max().item() <= 1e-5 and b.grad.abs().max().item() <= 1e-5
    # one-sided gradient requests and duplicate rows (clamp path)
    x, y = _case(MMD_CASES[3])
    x[5] = x[4]
    y[7] = x[4]
    xt = dev(x).requires_grad_(True)
    l = edrl_b200.MK_MMD(xt, dev(y), precision=name)
    l.backward()
    ref_l, _, dx, _ = O.mk_mmd_grad(x, y)
    assert np.isclose(l.item(), ref_l, rtol=ltol, atol=1e-6)
    assert np.abs(xt.grad.cpu().numpy() - dx).max() <= gtol * np.abs(dx).max()
    yt = dev(y).requires_grad_(True)
    l = edrl_b200.MK_MMD(dev(x), yt, precision=name)
    l.backward()
    _, _, _, dy = O.mk_mmd_grad(x, y)
    assert np.abs(yt.grad.cpu().numpy() - dy).max() <= gtol * np.abs(dy).max()
    # fp64 inputs: computed in fp32, cast back
    l64 = edrl_b200.MK_MMD(dev(x, torch.float64), dev(y, torch.float64), precision=name)
    assert l64.dtype == torch.float64 and np.isclose(l64.item(), ref_l, rtol=ltol, atol=1e-6)


def test_mk_mmd_errors():
    import edrl_b200
    with pytest.raises(RuntimeError):
        edrl_b200.MK_MMD(torch.randn(4, 8), torch.randn(4, 8))             # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        edrl_b200.MK_MMD(torch.randn(4, 8).cuda(), torch.randn(4, 9).cuda())
    with pytest.raises(ValueError):
        edrl_b200.MK_MMD(torch.randn(4, 8).cuda(), torch.randn(4, 8).cuda(), kernel_num=0)
    with pytest.raises(ValueError):
        edrl_b200.MK_MMD(torch.randn(4, 8).cuda(), torch.randn(4, 8).cuda(), precision="fp8")


# ---------------------------------------------------------------- C-ABI details: row ranges, tile shards
@pytest.mark.parametrize("flag", [0, 1])
def test_backward_row_ranges_and_tile_shards(raw, flag):
    ns, nt, d = 300, 212, 96
    rng = np.random.default_rng(3)
    x = rng.standard_normal((ns, d)).astype(np.float32)
    y = (rng.standard_normal((nt, d)) * 1.2 + 0.1).astype(np.float32)
    xd, yd = dev(x), dev(y)
    loss, stats, _, ws = raw.forward(xd, yd, flags=flag)
    full = raw.backward(ns, nt, d, stats, ws, 0, ns + nt, flags=flag)
    part = raw.backward(ns, nt, d, stats, ws, 131, 257, grad_out=1.0, flags=flag)
    torch.cuda.synchronize()
    np.testing.assert_allclose(part.cpu().numpy(), full[131:131 + 257].cpu().numpy(), rtol=0, atol=0)
    # upper-triangular tile list split over 3 "ranks": partial sums add up to the single-rank result
    tot = torch.zeros(2, dtype=torch.float64, device="cuda")
    for r in range(3):
        _, _, p, ws_r = raw.forward(xd, yd, flags=flag, tile_rank=r, tile_world=3)
        tot += p
    l2, s2 = raw.finalize(tot, ns, nt, ws_r)
    torch.cuda.synchronize()
    assert np.isclose(l2.item(), loss.item(), rtol=1e-6)
    np.testing.assert_allclose(s2.cpu().numpy()[:5], stats.cpu().numpy()[:5], rtol=1e-5)
    ref_l, _, dx, dy = O.mk_mmd_grad(x.astype(np.float64), y.astype(np.float64))
    assert np.isclose(loss.item(), ref_l, rtol=1e-3)


# ---------------------------------------------------------------- full-size properties (BASELINE config 2)
@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
def test_full_size_properties(mode):
    import edrl_b200
    name, flag, ltol, gtol = mode
    N, d = 8192, 512
    g = torch.Generator(device="cuda").manual_seed(1013)
    x = torch.randn(N, d, device="cuda", generator=g)
    y = torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1
    xt = x.clone().requires_grad_(True)
    yt = y.clone().requires_grad_(True)
    l = edrl_b200.MK_MMD(xt, yt, precision=name)
    l.backward()
    lv = l.item()
    assert np.isfinite(lv) and lv > 0
    # symmetry in the arguments
    assert np.isclose(edrl_b200.MK_MMD(y, x, precision=name).item(), lv, rtol=ltol)
    # scale covariance: the bandwidth is data derived, so (aX, aY) gives the same loss
    assert np.isclose(edrl_b200.MK_MMD(3.0 * x, 3.0 * y, precision=name).item(), lv, rtol=ltol * 3)
    # permutation invariance within each set
    px = x[torch.randperm(N, device="cuda")]
    assert np.isclose(edrl_b200.MK_MMD(px, y, precision=name).item(), lv, rtol=ltol)
    # translation invariance => gradients sum to zero over all rows
    gsum = (xt.grad.sum(0) + yt.grad.sum(0)).abs().max().item()
    gmax = max(xt.grad.abs().max().item(), yt.grad.abs().max().item())
    # (fp32 TMEM accumulation truncates, so the residual is a small systematic per-row bias)
    assert gsum / (2 * N) <= (1e-5 if name == "3xtf32" else 1e-4) * gmax, (gsum, gmax)
    # directional derivative against a forward difference (3xTF32 forward for the difference quotient)
    v = xt.grad / xt.grad.norm()
    t = 0.5
    lp = edrl_b200.MK_MMD(x + t * v, y, precision="3xtf32").item()
    lm = edrl_b200.MK_MMD(x - t * v, y, precision="3xtf32").item()
    fd = (lp - lm) / (2 * t)
    an = (xt.grad * v).sum().item()
    assert np.isclose(fd, an, rtol=5e-2), (fd, an)
    # against the strided-subsample oracle: compare with a smaller exact problem embedded? (not possible) ->
    # instead cross-check the two precision modes against each other
    l3 = edrl_b200.MK_MMD(x, y, precision="3xtf32").item()
    assert np.isclose(lv, l3, rtol=1e-3)


# ---------------------------------------------------------------- parity AT the quoted sizes (BASELINE configs[1], [3])
def _record(name, payload):
    """Measured errors of the full-size parity tests, for DESIGN.md (gpurun_out/ travels back from the GPU box)."""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "parity_full_size.json")
        try:
            with open(path) as f:
                data = json.load(f)
        except Exception:
            data = {}
        data[name] = payload
        with open(path, "w") as f:
            json.dump(data, f, indent=1)


@pytest.fixture(scope="module")
def reference_ops_8192():
    """The reference's own op sequence (oracle/cpu_port.mk_mmd_fwd_bwd: code/MMD.py:16-72 + autograd, the n x n
    matrices materialised) run on the GPU at N = 8192 per side, d = 512 -- in fp64 (the truth) and in fp32 (what the
    reference computes).  bench.py's inputs (seed 1013)."""
    from oracle import cpu_port
    N, d = 8192, 512
    g = torch.Generator(device="cuda").manual_seed(1013)
    x = torch.randn(N, d, device="cuda", generator=g)
    y = torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1
    l64, dx64, dy64 = cpu_port.mk_mmd_fwd_bwd(x.double(), y.double())
    torch.cuda.empty_cache()
    l32, dx32, dy32 = cpu_port.mk_mmd_fwd_bwd(x, y)
    torch.cuda.empty_cache()
    return x, y, (l64.item(), dx64, dy64), (l32.item(), dx32, dy32)


@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
def test_full_size_vs_reference_ops(reference_ops_8192, mode):
    """N = 8192 per side, d = 512 (the headline workload): loss and the FULL gradients of every precision mode against
    the reference op sequence evaluated in fp64 on the same GPU, at the tolerances of SURVEY.md 8c; and against the
    reference's fp32 evaluation at the reference's own error level."""
    import edrl_b200
    name, flag, ltol, gtol = mode
    x, y, (l64, dx64, dy64), (l32, dx32, dy32) = reference_ops_8192
    xt = x.clone().requires_grad_(True)
    yt = y.clone().requires_grad_(True)
    loss = edrl_b200.MK_MMD(xt, yt, precision=name)
    loss.backward()
    gmax = max(dx64.abs().max().item(), dy64.abs().max().item())
    lerr = abs(loss.item() - l64) / abs(l64)
    gerr = max((xt.grad.double() - dx64).abs().max().item(), (yt.grad.double() - dy64).abs().max().item()) / gmax
    ref_lerr = abs(l32 - l64) / abs(l64)
    ref_gerr = max((dx32.double() - dx64).abs().max().item(), (dy32.double() - dy64).abs().max().item()) / gmax
    _record(f"N8192_d512_{name}", {"loss": loss.item(), "loss_fp64_reference_ops": l64, "loss_rel_err": lerr,
                                   "grad_max_err_over_gmax": gerr, "reference_fp32_loss_rel_err": ref_lerr,
                                   "reference_fp32_grad_err_over_gmax": ref_gerr, "tolerance": [ltol, gtol]})
    assert lerr <= ltol, (loss.item(), l64)
    assert gerr <= gtol, gerr
    # against the reference's own fp32 numbers: both sit within their error of the fp64 truth
    assert abs(loss.item() - l32) <= (ltol + 2e-4) * abs(l32)
    assert max((xt.grad - dx32).abs().max().item(), (yt.grad - dy32).abs().max().item()) <= (gtol + 2e-4) * gmax


@pytest.fixture(scope="module")
def blockwise_65536():
    """fp64 row-blocked restatement of the reference (oracle/blockwise.py) at N = 65536 per side, d = 1024 (BASELINE
    configs[3], rank 0's generator seeds of bench.py: the whole problem on one GPU): loss, sigma_0 and the gradient rows
    of 256 sampled rows (both ends of each set, the panel boundaries, random ones)."""
    from oracle import blockwise
    N, d = 65536, 1024
    g = torch.Generator(device="cuda").manual_seed(2000)
    x = torch.randn(N, d, device="cuda", generator=g)
    y = torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1
    gr = torch.Generator().manual_seed(5)
    rows = torch.cat([torch.tensor([0, 1, 127, 128, N - 1, N, N + 1, 2 * N - 1, N + 255, N + 256]),
                      torch.randint(0, 2 * N, (246,), generator=gr)]).unique()
    loss, m, s0, grad = blockwise.mk_mmd_blockwise(x, y, rows=rows.cuda(), block=2048)
    torch.cuda.empty_cache()
    return x, y, rows.cuda(), loss.item(), s0.item(), grad


@pytest.mark.parametrize("mode", [MODES[1], MODES[3]], ids=lambda m: m[0])
def test_sharded_size_sampled_rows_vs_blockwise_oracle(blockwise_65536, mode):
    """N = 65536 per side, d = 1024 on ONE GPU (the quad sweep, the kernel behind every configs[3] number): loss,
    bandwidth and sampled gradient rows against the fp64 row-blocked oracle."""
    import edrl_b200
    name, flag, ltol, gtol = mode
    x, y, rows, l64, s0, g64 = blockwise_65536
    N = x.shape[0]
    xt = x.clone().requires_grad_(True)
    yt = y.clone().requires_grad_(True)
    loss = edrl_b200.MK_MMD(xt, yt, precision=name)
    loss.backward()
    gz = torch.cat([xt.grad, yt.grad])[rows].double()
    gmax = g64.abs().max().item()
    lerr = abs(loss.item() - l64) / abs(l64)
    gerr = (gz - g64).abs().max().item() / gmax
    _, st = edrl_b200.mk_mmd_with_stats(x, y, precision=name)
    _record(f"N65536_d1024_{name}", {"loss": loss.item(), "loss_fp64_blockwise": l64, "loss_rel_err": lerr,
                                     "sampled_rows": int(rows.numel()), "grad_max_err_over_gmax": gerr,
                                     "sigma0": st[1].item(), "sigma0_fp64": s0, "tolerance": [ltol, gtol]})
    assert lerr <= ltol, (loss.item(), l64)
    assert gerr <= gtol, gerr
    assert np.isclose(st[1].item(), s0, rtol=1e-5)
    del xt, yt
    torch.cuda.empty_cache()


@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
def test_midsize_vs_oracle(mode):
    """N=2048/side, d=512 -- the largest size the fp64 numpy oracle finishes in seconds."""
    import edrl_b200
    name, flag, ltol, gtol = mode
    x, y = mmd_inputs(1011, 2048, 2048, 512, 0.1, 1.25)
    x, y = x.numpy(), y.numpy()
    xt = dev(x).requires_grad_(True)
    yt = dev(y).requires_grad_(True)
    l = edrl_b200.MK_MMD(xt, yt, precision=name)
    l.backward()
    ref_l, _, dx, dy = O.mk_mmd_grad(x, y)
    assert np.isclose(l.item(), ref_l, rtol=ltol, atol=1e-6), (l.item(), ref_l)
    gmax = max(np.abs(dx).max(), np.abs(dy).max())
    assert np.abs(xt.grad.cpu().numpy() - dx).max() <= gtol * gmax
    assert np.abs(yt.grad.cpu().numpy() - dy).max() <= gtol * gmax


# ---------------------------------------------------------------- fused forward + gradient pass (TF32)
@pytest.mark.parametrize("ns,nt,d", [(300, 212, 96), (37, 53, 24), (1024, 1024, 512), (200, 200, 700)])
def test_fused_pass_matches_separate_kernels_and_oracle(raw, ns, nt, d):
    from gpu_util import raw_apply_grad, raw_forward_grad
    rng = np.random.default_rng(ns + d)
    x = rng.standard_normal((ns, d)).astype(np.float32)
    y = (rng.standard_normal((nt, d)) * 1.2 + 0.1).astype(np.float32)
    xd, yd = dev(x), dev(y)
    n = ns + nt
    # separate forward + tile-recomputing backward
    loss0, stats0, _, ws0 = raw.forward(xd, yd, flags=0)
    dz0 = raw.backward(ns, nt, d, stats0, ws0, 0, n, grad_out=1.5, flags=0)
    # fused
    loss1, stats1, _, u, ws1 = raw_forward_grad(raw, xd, yd, 0, n)
    dz1 = raw_apply_grad(raw, ns, nt, d, stats1, u, ws1, 0, n, grad_out=1.5)
    torch.cuda.synchronize()
    assert np.isclose(loss1.item(), loss0.item(), rtol=1e-5)
    np.testing.assert_allclose(stats1.cpu().numpy()[:5], stats0.cpu().numpy()[:5], rtol=1e-4, atol=1e-9)
    gmax = dz0.abs().max().item()
    # two TF32-level evaluations (the separate backward rounds G including c, the fused pass adds c in closed form)
    assert (dz1 - dz0).abs().max().item() <= 2e-3 * gmax, (dz1 - dz0).abs().max().item() / gmax
    ref, _, dx, dy = O.mk_mmd_grad(x.astype(np.float64), y.astype(np.float64), grad_out=1.5)
    assert np.isclose(loss1.item(), ref, rtol=1e-3, atol=1e-6)
    gm = max(np.abs(dx).max(), np.abs(dy).max())
    assert np.abs(dz1.cpu().numpy() - np.concatenate([dx, dy])).max() <= 2e-3 * gm
    # two row ranges, partial sums only (what one rank of a sharded evaluation runs), summed over 2 "ranks"
    h_s, h_t = ns // 2, nt // 2
    tot = torch.zeros(2, dtype=torch.float64, device="cuda")
    pieces = []
    for (r0, c0, r1, c1) in ((0, h_s, ns, h_t), (h_s, ns - h_s, ns + h_t, nt - h_t)):
        _, _, part, u2, ws2 = raw_forward_grad(raw, xd, yd, r0, c0, r1, c1, finalize=0)
        tot += part
        pieces.append((r0, c0, r1, c1, u2, ws2))
    l2, s2 = raw.finalize(tot, ns, nt, pieces[-1][5])
    torch.cuda.synchronize()
    assert np.isclose(l2.item(), loss1.item(), rtol=1e-6)
    # (a different row range has a different work list -- column sweeps split into a different number of slabs -- so
    #  the same fp32 terms are summed in a different order: agreement to a few fp32 ulps of |g|_inf, not bit-exact)
    for (r0, c0, r1, c1, u2, ws2) in pieces:
        dzp = raw_apply_grad(raw, ns, nt, d, s2, u2, ws2, r0, c0, r1, c1, grad_out=1.5)
        np.testing.assert_allclose(dzp[:c0].cpu().numpy(), dz1[r0:r0 + c0].cpu().numpy(), rtol=0, atol=2e-5 * gmax)
        np.testing.assert_allclose(dzp[c0:].cpu().numpy(), dz1[r1:r1 + c1].cpu().numpy(), rtol=0, atol=2e-5 * gmax)


@pytest.mark.parametrize("ns,nt,d", [(5000, 4600, 700), (3000, 2500, 1100), (9600, 9400, 300)])
@pytest.mark.parametrize("prec,flag", [("tf32", 0), ("tf32h", 2), ("f16s", 4), ("3xtf32", 1)])
def test_work_list_shapes_fused_vs_separate_kernels(raw, ns, nt, d, prec, flag):
    """Shapes whose work list mixes whole panels, column slabs and several 512-column feature passes (two passes at
    d = 700, three at 1100; 150 / 129 / 149 virtual panels on 74 SM pairs): the persistent fused sweep against the
    independent forward + tile-recomputing backward kernels, and two disjoint row ranges against the whole."""
    from gpu_util import raw_apply_grad, raw_forward_grad
    g = torch.Generator(device="cuda").manual_seed(ns + d)
    xd = torch.randn(ns, d, device="cuda", generator=g)
    yd = torch.randn(nt, d, device="cuda", generator=g) * 1.2 + 0.1
    n = ns + nt
    loss0, stats0, _, ws0 = raw.forward(xd, yd, flags=0)
    dz0 = raw.backward(ns, nt, d, stats0, ws0, 0, n, grad_out=1.5, flags=0)
    ws = raw.workspace(ns, nt, d, flag)
    loss1, stats1, _, u, ws1 = raw_forward_grad(raw, xd, yd, 0, n, ws=ws, flags=flag)
    dz1 = raw_apply_grad(raw, ns, nt, d, stats1, u, ws1, 0, n, grad_out=1.5, flags=flag)
    torch.cuda.synchronize()
    assert np.isclose(loss1.item(), loss0.item(), rtol=2e-5), (loss1.item(), loss0.item())
    gmax = dz0.abs().max().item()
    assert (dz1 - dz0).abs().max().item() <= 2e-3 * gmax, (dz1 - dz0).abs().max().item() / gmax
    # two row ranges (one rank of a 2-rank sharded evaluation) reproduce their rows of the whole
    h_s, h_t = ns // 2 + 17, nt // 2 - 5
    tot = torch.zeros(2, dtype=torch.float64, device="cuda")
    pieces = []
    for (r0, c0, r1, c1) in ((0, h_s, ns, h_t), (h_s, ns - h_s, ns + h_t, nt - h_t)):
        wsp = raw.workspace(ns, nt, d, flag)
        _, _, part, u2, ws2 = raw_forward_grad(raw, xd, yd, r0, c0, r1, c1, finalize=0, ws=wsp, flags=flag)
        tot += part
        pieces.append((r0, c0, r1, c1, u2, ws2))
    l2, s2 = raw.finalize(tot, ns, nt, pieces[-1][5])
    torch.cuda.synchronize()
    assert np.isclose(l2.item(), loss1.item(), rtol=1e-6)
    for (r0, c0, r1, c1, u2, ws2) in pieces:
        dzp = raw_apply_grad(raw, ns, nt, d, s2, u2, ws2, r0, c0, r1, c1, grad_out=1.5, flags=flag)
        assert (dzp[:c0] - dz1[r0:r0 + c0]).abs().max().item() <= 2e-5 * gmax
        assert (dzp[c0:] - dz1[r1:r1 + c1]).abs().max().item() <= 2e-5 * gmax


@pytest.mark.parametrize("prec,flag", [("tf32", 0), ("tf32h", 2), ("f16s", 4)])
@pytest.mark.parametrize("ns,nt,d", [(700, 650, 1100), (900, 1100, 400)])
def test_generic_bandwidths_in_the_sweeps(raw, ns, nt, d, prec, flag):
    """kernel_mul != 2 / kernel_num != 5 take the generic exponential path of the epilogue: the quad sweep (d = 1100) and
    the pair sweep (d = 400) against the separate forward + backward entry points and the fp64 oracle."""
    from gpu_util import raw_apply_grad, raw_forward_grad
    rng = np.random.default_rng(ns + d)
    x = rng.standard_normal((ns, d)).astype(np.float32)
    y = (rng.standard_normal((nt, d)) * 1.3 - 0.2).astype(np.float32)
    xd, yd = dev(x), dev(y)
    n = ns + nt
    for (mul, num) in ((1.5, 3), (3.0, 7)):
        loss0, stats0, _, ws0 = raw.forward(xd, yd, mul=mul, num=num, flags=0)
        dz0 = raw.backward(ns, nt, d, stats0, ws0, 0, n, grad_out=0.5, mul=mul, num=num, flags=0)
        ws = raw.workspace(ns, nt, d, flag)
        loss1, stats1, _, u, ws1 = raw_forward_grad(raw, xd, yd, 0, n, mul=mul, num=num, ws=ws, flags=flag)
        dz1 = raw_apply_grad(raw, ns, nt, d, stats1, u, ws1, 0, n, grad_out=0.5, flags=flag)
        torch.cuda.synchronize()
        assert np.isclose(loss1.item(), loss0.item(), rtol=2e-5), (mul, num)
        gmax = dz0.abs().max().item()
        assert (dz1 - dz0).abs().max().item() <= 2e-3 * gmax, (mul, num)
        ref, _, dx, dy = O.mk_mmd_grad(x.astype(np.float64), y.astype(np.float64), kernel_mul=mul, kernel_num=num,
                                       grad_out=0.5)
        assert np.isclose(loss1.item(), ref, rtol=1e-3, atol=1e-6), (mul, num)
        gm = max(np.abs(dx).max(), np.abs(dy).max())
        assert np.abs(dz1.cpu().numpy() - np.concatenate([dx, dy])).max() <= 2e-3 * gm, (mul, num)


# ---------------------------------------------------------------- host-side robustness: layouts, dtypes, streams, graphs
def test_noncontiguous_bf16_inputs_and_side_stream():
    import edrl_b200
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn(200, 2 * 96, device="cuda", generator=g)
    x = base[:, ::2]                                   # non-contiguous view
    y = torch.randn(150, 96, device="cuda", generator=g) * 1.2 + 0.1
    ref = edrl_b200.MK_MMD(x.contiguous(), y, precision="3xtf32").item()
    assert np.isclose(edrl_b200.MK_MMD(x, y, precision="3xtf32").item(), ref, rtol=1e-6)
    # bf16 inputs: computed in fp32 from the bf16 values, result cast back
    xb, yb = x.contiguous().bfloat16().requires_grad_(True), y.bfloat16().requires_grad_(True)
    lb = edrl_b200.MK_MMD(xb, yb)
    assert lb.dtype == torch.bfloat16
    lb.backward()
    assert xb.grad.dtype == torch.bfloat16 and torch.isfinite(xb.grad.float()).all()
    exact = edrl_b200.MK_MMD(xb.detach().float(), yb.detach().float(), precision="3xtf32").item()
    assert np.isclose(lb.float().item(), exact, rtol=2e-2)
    # work is enqueued on the current stream, whichever it is
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        xs = x.contiguous().requires_grad_(True)
        ls = edrl_b200.MK_MMD(xs, y)
        ls.backward()
    s.synchronize()
    xd = x.contiguous().requires_grad_(True)
    ld = edrl_b200.MK_MMD(xd, y)
    ld.backward()
    torch.cuda.synchronize()
    assert ls.item() == ld.item() and torch.equal(xs.grad, xd.grad)


def test_cuda_graph_capture_of_a_training_step():
    """The library only enqueues memsets and kernels on the caller's stream, so a fwd+bwd step is capturable."""
    import edrl_b200
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # static inputs + warm-up on the capture stream
        g = torch.Generator(device="cuda").manual_seed(9)
        xs = torch.randn(512, 128, device="cuda", generator=g)
        ys = torch.randn(384, 128, device="cuda", generator=g) * 1.3 + 0.2
        for _ in range(2):
            a = xs.clone().requires_grad_(True)
            b = ys.clone().requires_grad_(True)
            torch.autograd.grad(edrl_b200.MK_MMD(a, b), (a, b))
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a = xs.clone().requires_grad_(True)
        b = ys.clone().requires_grad_(True)
        l1 = edrl_b200.MK_MMD(a, b)
        gx1, gy1 = torch.autograd.grad(l1, (a, b))
    first = None
    for scale in (1.0, 1.5):                           # replay on new data in the captured input buffers
        xs.copy_(xs * scale)
        graph.replay()
        torch.cuda.synchronize()
        xe = xs.clone().requires_grad_(True)
        ye = ys.clone().requires_grad_(True)
        l2 = edrl_b200.MK_MMD(xe, ye)
        gx2, gy2 = torch.autograd.grad(l2, (xe, ye))
        assert np.isclose(l1.item(), l2.item(), rtol=1e-6)
        assert torch.allclose(gx1, gx2, rtol=1e-5, atol=1e-9) and torch.allclose(gy1, gy2, rtol=1e-5, atol=1e-9)
        if first is None:
            first = l1.item()
    assert not np.isclose(first, l1.item(), rtol=1e-3)


# ---------------------------------------------------------------- ragged shapes (tile / chunk / pass boundaries)
RAGGED = [(1, 1, 1), (1, 7, 3), (2, 2, 65), (127, 1, 63), (129, 127, 64), (128, 128, 33), (255, 257, 512),
          (100, 300, 513), (64, 64, 1025), (5, 9, 3072)]


@pytest.mark.parametrize("mode", MODES, ids=lambda m: m[0])
@pytest.mark.parametrize("ns,nt,d", RAGGED)
def test_ragged_shapes_vs_oracle(ns, nt, d, mode):
    import edrl_b200
    name, flag, ltol, gtol = mode
    rng = np.random.default_rng(ns * 7919 + nt * 31 + d)
    x = rng.standard_normal((ns, d)) * 0.7 - 0.1
    y = rng.standard_normal((nt, d)) * 1.1 + 0.3
    xt = dev(x).requires_grad_(True)
    yt = dev(y).requires_grad_(True)
    loss = edrl_b200.MK_MMD(xt, yt, precision=name)
    loss.backward()
    x32, y32 = xt.detach().cpu().numpy().astype(np.float64), yt.detach().cpu().numpy().astype(np.float64)
    ref, _, dx, dy = O.mk_mmd_grad(x32, y32)
    assert np.isclose(loss.item(), ref, rtol=ltol, atol=1e-6), (loss.item(), ref)
    gmax = max(np.abs(dx).max(), np.abs(dy).max())
    if ns + nt == 2:
        # one sample per side: the loss is a constant (scale covariance), the true gradient is identically zero and
        # what is left is rounding of the two cancelling terms -- bound it by the size of one term,
        # 4 |G| |x - y| <= 4 * 5 / sigma_0 * |x - y|
        _, st = edrl_b200.mk_mmd_with_stats(dev(x), dev(y), precision=name)
        gmax = 20.0 / st[1].item() * np.abs(x32 - y32).max()
    assert np.abs(xt.grad.cpu().numpy() - dx).max() <= gtol * gmax
    assert np.abs(yt.grad.cpu().numpy() - dy).max() <= gtol * gmax
    # loss-only evaluation (no gradient requested) takes the symmetric forward kernel
    with torch.no_grad():
        l0 = edrl_b200.MK_MMD(dev(x), dev(y), precision=name)
    assert np.isclose(l0.item(), ref, rtol=ltol, atol=1e-6)


def test_binary16_container_modes_are_numerically_the_tf32_result():
    """The binary16 copies carry the same 11-bit significands as the TF32 operands, so the modes agree far
    inside the TF32 tolerance (only round-half tie-breaking of G, the accumulation order inside the tensor core and
    binary16 underflow can differ)."""
    import edrl_b200
    x, y = mmd_inputs(1011, 1024, 768, 512, 0.1, 1.25)
    outs = {}
    for prec in ("tf32", "tf32h", "f16s"):
        xt = dev(x.numpy()).requires_grad_(True)
        yt = dev(y.numpy()).requires_grad_(True)
        l = edrl_b200.MK_MMD(xt, yt, precision=prec)
        l.backward()
        outs[prec] = (l.item(), xt.grad.clone(), yt.grad.clone())
    assert outs["tf32"][0] == outs["tf32h"][0]                     # the forward sums do not involve G.Z
    gmax = outs["tf32"][1].abs().max().item()
    assert (outs["tf32"][1] - outs["tf32h"][1]).abs().max().item() <= 2e-5 * gmax
    assert (outs["tf32"][2] - outs["tf32h"][2]).abs().max().item() <= 2e-5 * gmax
    # F16S: the Gram operands are the same values too; only the fp32 accumulation order inside the MMA differs
    assert abs(outs["tf32"][0] - outs["f16s"][0]) <= 2e-6 * abs(outs["tf32"][0])
    assert (outs["tf32"][1] - outs["f16s"][1]).abs().max().item() <= 5e-5 * gmax
    assert (outs["tf32"][2] - outs["f16s"][2]).abs().max().item() <= 5e-5 * gmax
    # badly scaled columns (1e-6 .. 1e6) exercise the per-column binary16 scale (and, for F16S, binary16 underflow
    # of the small columns in the Gram operand -- they do not contribute to an fp32 distance anyway)
    scale = torch.logspace(-6, 6, 512, dtype=torch.float64)
    xs, ys = (x * scale).numpy(), (y * scale).numpy()
    ref, _, dxr, _ = O.mk_mmd_grad(xs, ys)
    for prec in ("tf32h", "f16s"):
        xt = dev(xs).requires_grad_(True)
        l = edrl_b200.MK_MMD(xt, dev(ys), precision=prec)
        l.backward()
        assert np.isclose(l.item(), ref, rtol=1e-3), prec
        assert np.abs(xt.grad.cpu().numpy() - dxr).max() <= 2e-3 * np.abs(dxr).max(), prec


_HYBRID_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import edrl_b200
ns, nt, d = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
g = torch.Generator(device="cuda").manual_seed(99)
x = torch.randn(ns, d, device="cuda", generator=g, requires_grad=True)
y = (torch.randn(nt, d, device="cuda", generator=g) * 1.2 + 0.05).requires_grad_(True)
loss = edrl_b200.MK_MMD(x, y, precision="tf32")
loss.backward()
out = dict(loss=loss.item(), gx=x.grad.cpu().numpy(), gy=y.grad.cpu().numpy())
# the same step captured in a CUDA graph: the forked stream of the hybrid launch has to join the capture
xs, ys = x.detach().clone().requires_grad_(True), y.detach().clone().requires_grad_(True)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        xs.grad = None; ys.grad = None
        edrl_b200.MK_MMD(xs, ys, precision="tf32").backward()
torch.cuda.current_stream().wait_stream(side)
cg = torch.cuda.CUDAGraph()
xs.grad.zero_(); ys.grad.zero_()
with torch.cuda.graph(cg):
    l2 = edrl_b200.MK_MMD(xs, ys, precision="tf32")
    l2.backward()
xs.grad.zero_()
cg.replay()
torch.cuda.synchronize()
out["graph_loss"] = l2.item()
out["graph_gx"] = xs.grad.cpu().numpy()
# two row ranges in one call (a rank's source rows and target rows): the row-block sharded entry on a 1-rank group
import os, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 300))
dist.init_process_group("nccl", rank=0, world_size=1)
x2, y2 = x.detach().clone().requires_grad_(True), y.detach().clone().requires_grad_(True)
l3 = edrl_b200.sharded_MK_MMD(x2, y2, precision="tf32")
l3.backward()
torch.cuda.synchronize()
out["two_range_loss"] = l3.item()
out["two_range_gx"] = x2.grad.cpu().numpy()
out["two_range_gy"] = y2.grad.cpu().numpy()
dist.destroy_process_group()
np.savez(sys.argv[2], **out)
"""


@pytest.mark.parametrize("ns,nt,d", [(1500, 1400, 1024), (700, 650, 1100)])
def test_hybrid_quad_plus_pair_launch_matches_the_single_launch(tmp_path, ns, nt, d):
    """d > 768: the last row panels go to a pair kernel on a forked stream next to the 4-CTA-cluster kernel
    (csrc/mmd.cu make_hybrid).  Forced here on a small shape (EDRL_MMD_HYBRID=2 is read once per process, hence the child
    processes) and compared with the single launch (=0) and the fp64 oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    # "2r": most panels to the pair kernel (its part then starts before the second row range does)
    for mode in ("0", "2", "2r"):
        path = str(tmp_path / f"h{mode}.npz")
        env = dict(os.environ, EDRL_MMD_HYBRID=mode[0])
        if mode == "2r":
            env["EDRL_MMD_HYBRID_RATIO"] = "0.05"
        r = subprocess.run([sys.executable, "-c", _HYBRID_CHILD, root, path, str(ns), str(nt), str(d)], env=env,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[mode] = np.load(path)
    a, b, c = res["0"], res["2"], res["2r"]
    assert abs(float(c["loss"]) - float(a["loss"])) <= 2e-6 * max(1.0, abs(float(a["loss"])))
    assert np.abs(c["gx"] - a["gx"]).max() <= 2e-5 * float(np.abs(a["gx"]).max())
    assert np.abs(c["gy"] - a["gy"]).max() <= 2e-5 * float(np.abs(a["gy"]).max())
    gmax = float(np.abs(a["gx"]).max())
    assert abs(float(a["loss"]) - float(b["loss"])) <= 2e-6 * max(1.0, abs(float(a["loss"])))
    assert np.abs(a["gx"] - b["gx"]).max() <= 2e-5 * gmax          # same products, different summation order of the slabs
    assert np.abs(a["gy"] - b["gy"]).max() <= 2e-5 * float(np.abs(a["gy"]).max())
    assert abs(float(b["graph_loss"]) - float(b["loss"])) <= 1e-6
    assert np.abs(b["graph_gx"] - b["gx"]).max() <= 2e-5 * gmax
    # the two-range call (source rows, then target rows: the second range starts in the middle of the panel list)
    for r in (a, b, c):
        assert abs(float(r["two_range_loss"]) - float(a["loss"])) <= 2e-6 * max(1.0, abs(float(a["loss"])))
        assert np.abs(r["two_range_gx"] - a["gx"]).max() <= 2e-5 * gmax
        assert np.abs(r["two_range_gy"] - a["gy"]).max() <= 2e-5 * float(np.abs(a["gy"]).max())
    # and against the fp64 oracle
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn(ns, d, device="cuda", generator=g)
    y = torch.randn(nt, d, device="cuda", generator=g) * 1.2 + 0.05
    want, _, gx, gy = O.mk_mmd_grad(x.cpu().numpy(), y.cpu().numpy())
    tol = dict((m[0], m) for m in MODES)["tf32"]
    assert np.isclose(float(b["loss"]), want, rtol=tol[2], atol=1e-6)
    assert np.abs(b["gx"] - gx).max() <= tol[3] * max(np.abs(gx).max(), np.abs(gy).max())


def test_separate_backward_is_the_first_call_of_the_autograd_thread(tmp_path):
    """`EDRL_MMD_FUSED=0` (and 3xTF32 at d > 768): `edrl_mmd_backward` builds tensor maps, and in a fresh process it is the
    first call this library sees on torch's autograd thread -- which has a device selected but possibly no context bound
    (cuTensorMapEncodeTiled: CUDA_ERROR_INVALID_CONTEXT before the device guard bound one)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "import edrl_b200\n"
        "for prec, d in (('tf32', 8), ('3xtf32', 8), ('3xtf32', 800)):\n"
        "    x = torch.randn(6, d, device='cuda', requires_grad=True)\n"
        "    y = torch.randn(5, d, device='cuda', requires_grad=True)\n"
        "    edrl_b200.MK_MMD(x, y, precision=prec).backward()\n"
        "    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, EDRL_MMD_FUSED="0"), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("ns,nt,d", [(700, 650, 1100), (300, 280, 3072)])
def test_3xtf32_beyond_768_columns_vs_oracle(ns, nt, d):
    """3xTF32 at d > 768 (the reference's own feature width is 3072): the fused pair sweep, one Gram pass per 512 feature
    columns, at the fp32-level tolerances of the mode -- forward + backward through the public API, and the separate
    forward / backward entry points (EDRL_MMD_FUSED=0 takes those) through the C-ABI."""
    import edrl_b200
    from gpu_util import RawMMD
    name, flag, ltol, gtol = MODES[0] if MODES[0][0] == "3xtf32" else [m for m in MODES if m[0] == "3xtf32"][0]
    rng = np.random.default_rng(ns + d)
    x = rng.standard_normal((ns, d)).astype(np.float32)
    y = (rng.standard_normal((nt, d)) * 1.2 + 0.1).astype(np.float32)
    xt, yt = dev(x).requires_grad_(True), dev(y).requires_grad_(True)
    loss = edrl_b200.MK_MMD(xt, yt, precision="3xtf32")
    loss.backward()
    ref, _, dx, dy = O.mk_mmd_grad(x.astype(np.float64), y.astype(np.float64))
    assert np.isclose(loss.item(), ref, rtol=ltol, atol=1e-6), (loss.item(), ref)
    gm = max(np.abs(dx).max(), np.abs(dy).max())
    assert np.abs(xt.grad.cpu().numpy() - dx).max() <= gtol * gm
    assert np.abs(yt.grad.cpu().numpy() - dy).max() <= gtol * gm
    raw = RawMMD()
    loss0, stats0, _, ws0 = raw.forward(xt.detach(), yt.detach(), flags=flag)
    dz0 = raw.backward(ns, nt, d, stats0, ws0, 0, ns + nt, grad_out=1.0, flags=flag)
    torch.cuda.synchronize()
    assert np.isclose(loss0.item(), ref, rtol=ltol, atol=1e-6)
    assert np.abs(dz0.cpu().numpy() - np.concatenate([dx, dy])).max() <= gtol * gm
