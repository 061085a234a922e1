"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python oracle/gen_golden.py

The reference has no tests or golden vectors for this path,
so these fixtures -- outputs of the reference's own ``MK_MMD`` (code/MMD.py:46) and
``EPRL.forward`` (code/fusion_net.py:133) under torch CPU -- are the pin for the
numpy oracle and, through it, for the CUDA kernels.  The reference modules are
imported unmodified; ``EPRL.gaussian_noise`` and the encoder output are *recorded*
through instance-level wrappers / hooks (no behaviour change).

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (seed, n_s, n_t, d, shift, scale) -- SURVEY.md section 4 known-answer table
MMD_CASES = [
    (101, 4, 4, 8, 0.0, 1.0),
    (102, 32, 32, 3072, 0.1, 1.25),
    (103, 64, 64, 3072, 0.05, 1.1),
    (104, 37, 53, 24, 0.2, 1.3),
    (105, 256, 256, 512, 0.1, 1.25),
    (106, 1024, 1024, 512, 0.1, 1.25),
]
# extra (kernel_mul, kernel_num) variants on a small case: general-bandwidth path
MMD_VARIANTS = [(2.0, 5), (2.0, 3), (3.0, 4), (1.5, 7), (2.0, 1)]


def mmd_inputs(seed, ns, nt, d, shift, scale, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(ns, d, generator=g, dtype=torch.float64)
    y = torch.randn(nt, d, generator=g, dtype=torch.float64) * scale + shift
    return x.to(dtype), y.to(dtype)


def run_ref_mmd(mmd, x, y, mul=2.0, num=5):
    x = x.clone().requires_grad_(True)
    y = y.clone().requires_grad_(True)
    loss = mmd.MK_MMD(x, y, kernel_mul=mul, kernel_num=num)
    loss.backward()
    return loss.detach(), x.grad, y.grad


def gen_mmd():
    mmd = ref_loader.load_reference_mmd()
    out = {}
    for (seed, ns, nt, d, shift, scale) in MMD_CASES:
        x64, y64 = mmd_inputs(seed, ns, nt, d, shift, scale)
        l64, gx64, gy64 = run_ref_mmd(mmd, x64, y64)
        l32, gx32, gy32 = run_ref_mmd(mmd, x64.float(), y64.float())
        key = f"s{seed}"
        out[key + "_cfg"] = np.array([seed, ns, nt, d, shift, scale], np.float64)
        out[key + "_loss64"] = l64.numpy()
        out[key + "_loss32"] = l32.numpy()
        out[key + "_sumabs_gx64"] = gx64.abs().sum().numpy()
        out[key + "_sumabs_gy64"] = gy64.abs().sum().numpy()
        out[key + "_maxabs_g64"] = np.array(max(gx64.abs().max().item(), gy64.abs().max().item()))
        # strided subsample of the fp64 gradients (keeps the fixture small)
        rs = max(1, ns // 8)
        cs = max(1, d // 16)
        out[key + "_gx64_sub"] = gx64[::rs, ::cs].numpy()
        out[key + "_gy64_sub"] = gy64[::max(1, nt // 8), ::cs].numpy()
        out[key + "_sub_strides"] = np.array([rs, max(1, nt // 8), cs])
        if ns * d <= 4096:
            out[key + "_x"] = x64.numpy()
            out[key + "_y"] = y64.numpy()
            out[key + "_gx64"] = gx64.numpy()
            out[key + "_gy64"] = gy64.numpy()
        print(f"mmd seed={seed}: loss64={l64.item():.12e} loss32={l32.item():.8e} "
              f"sum|gx|={gx64.abs().sum().item():.10e} sum|gy|={gy64.abs().sum().item():.10e}")
    # hand-checkable case
    x = torch.tensor([[0.0], [1.0]], dtype=torch.float64)
    y = torch.tensor([[2.0], [3.0]], dtype=torch.float64)
    out["hand_kernel_row0"] = mmd.gaussian_kernel(x, y)[0].numpy()
    out["hand_loss"] = mmd.MK_MMD(x, y).numpy()
    # identical inputs -> loss 0, zero grad
    xi, _ = mmd_inputs(7, 6, 6, 5, 0, 1)
    l, gx, gy = run_ref_mmd(mmd, xi, xi.clone())
    out["ident_loss"] = l.numpy()
    out["ident_gmax"] = np.array(max(gx.abs().max().item(), gy.abs().max().item()))
    # (kernel_mul, kernel_num) variants on case 104
    x64, y64 = mmd_inputs(*MMD_CASES[3])
    for (mul, num) in MMD_VARIANTS:
        l, gx, gy = run_ref_mmd(mmd, x64, y64, mul, num)
        key = f"var_m{mul}_k{num}"
        out[key + "_loss"] = l.numpy()
        out[key + "_gx"] = gx.numpy()
        out[key + "_gy"] = gy.numpy()
    np.savez_compressed(os.path.join(GOLD, "mmd_reference.npz"), **out)


def _record_eprl(model, x, y):
    """Run the reference EPRL once, recording eps and the encoder output."""
    rec = {}
    orig_noise = model.gaussian_noise

    def noise(*a, **k):
        e = orig_noise(*a, **k)
        rec["eps"] = e.detach().clone()
        return e

    model.gaussian_noise = noise

    def enc_hook(_m, _inp, outp):
        rec["z"] = outp.detach().clone()
        if outp.requires_grad:
            outp.register_hook(lambda g: rec.__setitem__("dz", g.detach().clone()))

    h = model.encoder.register_forward_hook(enc_hook)
    try:
        with ref_loader.cuda_identity_if_no_gpu():
            outs = model(x, y)
    finally:
        h.remove()
        del model.gaussian_noise
    return outs, rec


def gen_eprl():
    fn = ref_loader.load_reference_fusion_net()
    out = {}
    cases = [
        # name, B, T, x_dim, z_dim, S, seed
        ("small", 4, 10, 12, 16, 128, 11),
        ("tok144", 3, 144, 20, 16, 112, 12),
        ("tok216", 5, 216, 24, 32, 100, 13),
    ]
    for (name, b, t, xd, zd, s, seed) in cases:
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            torch.manual_seed(seed)
            model = fn.EPRL(xd, z_dim=zd, sample_num=s, num_classes=2, seed=1, batch_size=b).to(dtype)
            x = torch.randn(b, t, xd, dtype=dtype)
            y = torch.randint(0, 2, (b,))
            key = f"{name}_{tag}"
            # ---- train branch ----
            model.train()
            torch.manual_seed(seed + 1000)
            # torch.normal(zeros, ones) draws float32 noise regardless of the module dtype
            (mu, sigma, loss, ztopk), rec = _record_eprl(model, x, y)
            model.zero_grad()
            loss.backward()
            out[key + "_cfg"] = np.array([b, t, xd, zd, s, seed])
            out[key + "_y"] = y.numpy()
            out[key + "_z"] = rec["z"].numpy()
            out[key + "_eps"] = rec["eps"].numpy()
            out[key + "_proxies"] = model.proxies.detach().numpy()
            out[key + "_mu"] = mu.detach().numpy()
            out[key + "_sigma"] = sigma.detach().numpy()
            out[key + "_loss"] = loss.detach().numpy()
            out[key + "_dz"] = rec["dz"].numpy()
            out[key + "_dproxies"] = model.proxies.grad.numpy()
            assert torch.equal(ztopk, rec["z"])          # z_topk = z (code/fusion_net.py:253)
            print(f"eprl {key} train loss={loss.item():.10e}")
            # ---- eval branch (T must be 144 or anything else -> mlp_3d needs 216) ----
            if t in (144, 216):
                model.eval()
                with torch.no_grad():
                    try:
                        (mu_e, sig_e, loss_e, z_e, ent_e), rec_e = _record_eprl(model, x, y)
                        out[key + "_eval_ok"] = np.array(1)
                        out[key + "_eval_z"] = rec_e["z"].numpy()
                        out[key + "_eval_eps"] = rec_e["eps"].numpy()
                        out[key + "_eval_loss"] = loss_e.numpy()
                        out[key + "_eval_entropy"] = ent_e.numpy()
                        mlp = model.mlp_2d if t == 144 else model.mlp_3d
                        out[key + "_eval_mlp_w"] = mlp[1].weight.detach().numpy()
                        out[key + "_eval_mlp_b"] = mlp[1].bias.detach().numpy()
                        out[key + "_eval_alpha"] = model.alpha.detach().numpy()
                        print(f"eprl {key} eval loss={loss_e.item():.10e} entropy={ent_e.item():.10e}")
                    except (IndexError, RuntimeError) as exc:   # the fragile mask indexing (:191)
                        out[key + "_eval_ok"] = np.array(0)
                        print(f"eprl {key} eval raised {type(exc).__name__}")
    # state_dict key names (checkpoint compatibility)
    model = fn.EPRL(1024, num_classes=2, sample_num=800, batch_size=4)
    names = sorted(model.state_dict().keys())
    out["state_dict_keys"] = np.array(names)
    out["state_dict_shapes"] = np.array([str(tuple(model.state_dict()[k].shape)) for k in names])
    np.savez_compressed(os.path.join(GOLD, "eprl_reference.npz"), **out)


def gen_dilr():
    """DILR.bt_loss_cross (code/fusion_net.py:656-677) of the unmodified reference, train-mode BatchNorm (batch statistics,
    running-stat update), forward values and gradients w.r.t. both inputs of a weighted sum of the six outputs."""
    import types
    fn = ref_loader.load_reference_fusion_net()
    out = {}
    # name, B, D (BatchNorm width), common_dim, batch_size (the divisor's), seed
    cases = [("tiny", 6, 16, 8, 6, 21), ("odd", 5, 24, 10, 8, 22), ("ref2048", 8, 2048, 1024, 8, 23)]
    w = np.array([1.0, 0.3, -0.2, 0.7, 0.1, 0.5])
    for (name, b, d, dc, bs, seed) in cases:
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            torch.manual_seed(seed)
            holder = types.SimpleNamespace(args=types.SimpleNamespace(batch_size=bs),
                                           bn1=torch.nn.BatchNorm1d(d, affine=False).to(dtype),
                                           bn2=torch.nn.BatchNorm1d(d, affine=False).to(dtype))
            holder.bn1.train()
            holder.bn2.train()
            z1 = (torch.randn(b, d, dtype=dtype) * 1.5 + 0.3).requires_grad_(True)
            z2 = (0.6 * z1.detach() + 0.8 * torch.randn(b, d, dtype=dtype) - 0.1).requires_grad_(True)
            vals = fn.DILR.bt_loss_cross(holder, z1, z2, dc)          # the reference's own method on a stand-in self
            total = sum(float(wi) * v for wi, v in zip(w, vals))
            total.backward()
            key = f"{name}_{tag}"
            out[key + "_cfg"] = np.array([b, d, dc, bs, seed])
            out[key + "_z1"] = z1.detach().numpy()
            out[key + "_z2"] = z2.detach().numpy()
            out[key + "_out"] = np.array([v.item() for v in vals])
            out[key + "_dz1"] = z1.grad.numpy()
            out[key + "_dz2"] = z2.grad.numpy()
            out[key + "_rm1"] = holder.bn1.running_mean.numpy()
            out[key + "_rv1"] = holder.bn1.running_var.numpy()
            print(f"dilr {key} loss_c={vals[0].item():.8e} loss_u={vals[3].item():.8e}")
    out["weights"] = w
    np.savez_compressed(os.path.join(GOLD, "dilr_reference.npz"), **out)


if __name__ == "__main__":
    if not ref_loader.reference_available():
        raise SystemExit("reference not mounted at " + ref_loader.REFERENCE_ROOT)
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if "--only-dilr" not in sys.argv:
        gen_mmd()
        gen_eprl()
    gen_dilr()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)), "bytes")
