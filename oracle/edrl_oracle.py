"""CPU oracle for the EDRL hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm.  It is the checker
for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (the package next to this directory) never imports anything from
``oracle/`` and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the unmodified
reference itself (``/root/reference/code/MMD.py`` and ``fusion_net.EPRL``
imported in the build container by ``oracle/gen_golden.py``) that are committed
under ``tests/golden/`` -- see ``tests/test_oracle_golden.py``.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "gaussian_kernel", "mk_mmd", "mk_mmd_grad", "mmd_bandwidth",
    "eprl_sample_proxies", "eprl_scores", "eprl_scores_hoisted", "eprl_split",
    "topk_rows", "eprl_proxy_loss", "eprl_train_forward", "eprl_train_backward",
    "eprl_eval_forward", "gather_rows", "select_gather", "dilr_bt_loss_cross",
]


# --------------------------------------------------------------------------
# Part A: multi-bandwidth Gaussian MMD
# --------------------------------------------------------------------------
def _l2_distance(total: np.ndarray) -> np.ndarray:
    """code/MMD.py:25-27 -- r_i + r_j - 2 z_i.z_j, clamped at 0."""
    sq = np.sum(total ** 2, axis=1, keepdims=True)
    l2 = sq + sq.T - 2.0 * (total @ total.T)
    return l2


def mmd_bandwidth(total: np.ndarray, kernel_mul: float, kernel_num: int) -> float:
    """code/MMD.py:31-34 -- sigma_0 = sum(L) / (n^2 - n) / mul^(num // 2)."""
    n = total.shape[0]
    l2 = np.maximum(_l2_distance(total), 0.0)
    return float(l2.sum() / (n * n - n) / (kernel_mul ** (kernel_num // 2)))


def gaussian_kernel(source, target, kernel_mul=2.0, kernel_num=5):
    """code/MMD.py:3-44.  Returns the [n, n] summed kernel matrix in the input dtype."""
    source = np.asarray(source)
    target = np.asarray(target)
    dt = source.dtype
    n = source.shape[0] + target.shape[0]
    total = np.concatenate([source, target], axis=0)          # :21
    l2 = np.maximum(_l2_distance(total), dt.type(0))          # :25-27
    length_scale = l2.sum(dtype=dt) / dt.type(n * n - n)      # :31
    length_scale = length_scale / dt.type(kernel_mul ** (kernel_num // 2))  # :34
    out = np.zeros_like(l2)
    for i in range(kernel_num):                               # :37-42
        out += np.exp(-l2 / (length_scale * dt.type(kernel_mul ** i)))
    return out


def mk_mmd(source, target, kernel_mul=2.0, kernel_num=5):
    """code/MMD.py:46-74.  |XX + YY - XY - YX| with the diagonal included."""
    source = np.asarray(source)
    target = np.asarray(target)
    k = gaussian_kernel(source, target, kernel_mul, kernel_num)
    ns, nt = source.shape[0], target.shape[0]
    dt = k.dtype.type
    xx = k[:ns, :ns].sum() / dt(ns * ns)                      # :66
    yy = k[ns:, ns:].sum() / dt(nt * nt)                      # :67
    xy = k[:ns, ns:].sum() / dt(ns * nt)                      # :68
    yx = k[ns:, :ns].sum() / dt(ns * nt)                      # :69
    return np.abs(xx + yy - xy - yx)                          # :72


def mk_mmd_grad(source, target, kernel_mul=2.0, kernel_num=5, grad_out=1.0):
    """Closed-form backward of code/MMD.py:46-74 (what autograd does to it).

    The bandwidth is NOT detached in the reference (code/MMD.py:31-37), which adds
    the uniform term ``c``; ``clamp(min=0)`` back-propagates through ``L_raw >= 0``.
    Returns (loss, signed_mean M, dX, dY).  SURVEY.md section 8a row A6.
    """
    x = np.asarray(source)
    y = np.asarray(target)
    dt = x.dtype
    ns, nt = x.shape[0], y.shape[0]
    n = ns + nt
    z = np.concatenate([x, y], axis=0)
    l_raw = _l2_distance(z)
    l2 = np.maximum(l_raw, 0.0)
    mask = (l_raw >= 0.0).astype(dt)
    half = kernel_mul ** (kernel_num // 2)
    sigma0 = l2.sum() / (n * n - n) / half
    a = np.concatenate([np.full(ns, 1.0 / ns), np.full(nt, -1.0 / nt)]).astype(dt)
    w = np.outer(a, a)
    ksum = np.zeros_like(l2)
    amat = np.zeros_like(l2)
    dmat = np.zeros_like(l2)
    for k in range(kernel_num):
        sk = sigma0 * kernel_mul ** k
        e = np.exp(-l2 / sk)
        ksum += e
        amat -= e / sk
        dmat += e * l2 / (sk * sigma0)
    # block means exactly as code/MMD.py:66-72 (so identical inputs give M == 0 exactly)
    m = float(ksum[:ns, :ns].sum() / (ns * ns) + ksum[ns:, ns:].sum() / (nt * nt)
              - ksum[:ns, ns:].sum() / (ns * nt) - ksum[ns:, :ns].sum() / (ns * nt))
    dsum = float((w * dmat).sum())
    c = dsum / ((n * n - n) * half)
    g = (w * amat + c) * mask
    sgn = np.sign(m)
    dz = grad_out * sgn * 4.0 * (g.sum(axis=1, keepdims=True) * z - g @ z)
    return abs(m), m, dz[:ns].astype(dt), dz[ns:].astype(dt)


# --------------------------------------------------------------------------
# Part B: Essence-Point scoring and top-k selection (EPRL)
# --------------------------------------------------------------------------
_EPS = 1e-12  # torch.nn.functional.normalize default eps


def eprl_sample_proxies(mu, sigma, eps_noise):
    """code/fusion_net.py:143-146 -- z_p = mu[:,None,:] + sigma[:,None,:] * eps."""
    return mu[:, None, :] + sigma[:, None, :] * eps_noise


def _normalize_dim1(x):
    """F.normalize(x, dim=1) for a 3-D tensor (code/fusion_net.py:149-150)."""
    nrm = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
    return x / np.maximum(nrm, x.dtype.type(_EPS))


def eprl_scores(z, mu, sigma, eps_noise):
    """code/fusion_net.py:143-150,221-225 followed literally.

    z [B,T,F]; mu, sigma [C,F]; eps_noise [C,S,F]  ->  att [B,C,S].
    Materialises [B,C,T,S] like the reference: small cases only.
    """
    z_p = eprl_sample_proxies(mu, sigma, eps_noise)
    z_n = _normalize_dim1(z)            # over TOKENS (dim=1 of [B,T,F])
    z_pn = _normalize_dim1(z_p)         # over SAMPLES (dim=1 of [C,S,F])
    att = np.einsum("btf,csf->bcts", z_n, z_pn)   # :223
    return att.mean(axis=2)             # :224-225 (permute + mean over tokens)


def eprl_scores_hoisted(z, mu, sigma, eps_noise):
    """Same quantity with the token mean hoisted in front of the contraction
    (SURVEY.md section 8a row B4): att = zbar . z_pn, zbar = sum_t z / (T * max(|z|, eps))."""
    z_p = eprl_sample_proxies(mu, sigma, eps_noise)
    z_pn = _normalize_dim1(z_p)
    t = z.shape[1]
    nrm = np.sqrt(np.sum(z * z, axis=1))
    zbar = z.sum(axis=1) / (t * np.maximum(nrm, z.dtype.type(_EPS)))
    return np.einsum("bf,csf->bcs", zbar, z_pn)


def eprl_split(att, y):
    """code/fusion_net.py:227-234 -- positives = the label's class row, negatives =
    the remaining classes concatenated in class-major order."""
    b, c, s = att.shape
    y = np.asarray(y).astype(np.int64)
    if np.any((y < 0) | (y > 1)):
        # proxies_dict = {"0": 0, "1": 1}  (code/fusion_net.py:101)
        raise KeyError("label outside the reference's two-class proxies_dict")
    pos = att[np.arange(b), y, :]
    neg = np.stack([np.concatenate([att[i, cc] for cc in range(c) if cc != y[i]])
                    for i in range(b)]) if c > 1 else np.zeros((b, 0), att.dtype)
    return pos, neg


def topk_rows(x, k):
    """torch.topk(x, k, dim=1) (code/fusion_net.py:236-238): the k largest per row,
    sorted descending; ties resolved lowest-index-first (the rule the CUDA select
    kernel implements; torch leaves tie order unspecified)."""
    x = np.asarray(x)
    if k > x.shape[1]:
        raise RuntimeError("selected index k out of range")
    order = np.argsort(-x, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(x, order, axis=1)
    return vals, order.astype(np.int32)


def eprl_proxy_loss(top_pos, top_neg):
    """code/fusion_net.py:240-243 -- mean_b exp(-mean(top_pos_b) + mean(top_neg_b))."""
    return np.mean(np.exp(-top_pos.mean(axis=1) + top_neg.mean(axis=1)))


def eprl_train_forward(z, mu, sigma, eps_noise, y, k=100):
    """code/fusion_net.py:137-150,220-243: proxy_loss and the intermediates."""
    att = eprl_scores(z, mu, sigma, eps_noise)
    pos, neg = eprl_split(att, y)
    tp, ip = topk_rows(pos, k)
    tn, in_ = topk_rows(neg, k)
    loss = eprl_proxy_loss(tp, tn)
    return dict(att=att, pos_val=tp, pos_idx=ip, neg_val=tn, neg_idx=in_, loss=loss)


def eprl_train_backward(z, mu, sigma, eps_noise, y, k=100, grad_out=1.0):
    """Closed-form backward of the train branch (what autograd does to
    code/fusion_net.py:137-150,220-243).  Returns dict(dz, dmu, dsigma)."""
    dt = z.dtype
    b, t, f = z.shape
    c, s, _ = eps_noise.shape
    fw = eprl_train_forward(z, mu, sigma, eps_noise, y, k)
    e = np.exp(-fw["pos_val"].mean(axis=1) + fw["neg_val"].mean(axis=1))      # [B]
    coef = grad_out * e / (b * k)
    datt = np.zeros((b, c, s), dt)
    yy = np.asarray(y).astype(np.int64)
    for i in range(b):
        datt[i, yy[i], fw["pos_idx"][i]] += -coef[i]
        others = [cc for cc in range(c) if cc != yy[i]]
        for j in fw["neg_idx"][i]:
            datt[i, others[j // s], j % s] += coef[i]
    # score = zbar . z_pn
    z_p = eprl_sample_proxies(mu, sigma, eps_noise)
    nrm_p = np.sqrt(np.sum(z_p * z_p, axis=1))                # [C,F]
    q = np.maximum(nrm_p, dt.type(_EPS))
    z_pn = z_p / q[:, None, :]
    nrm = np.sqrt(np.sum(z * z, axis=1))                      # [B,F]
    m = np.maximum(nrm, dt.type(_EPS))
    ssum = z.sum(axis=1)
    zbar = ssum / (t * m)
    dzbar = np.einsum("bcs,csf->bf", datt, z_pn)
    dz_pn = np.einsum("bcs,bf->csf", datt, zbar)
    # token statistics backward
    dm = -dzbar * ssum / (t * m * m)
    live = (nrm > _EPS).astype(dt)
    dz = (dzbar / (t * m))[:, None, :] + (dm * live / np.where(nrm > 0, nrm, 1))[:, None, :] * z
    # proxy normalisation backward
    proj = np.sum(dz_pn * z_p, axis=1)                        # [C,F]
    live_p = (nrm_p > _EPS).astype(dt)
    dz_p = dz_pn / q[:, None, :] - z_p * (proj * live_p / (q * q * np.where(nrm_p > 0, nrm_p, 1)))[:, None, :]
    dmu = dz_p.sum(axis=1)
    dsigma = (dz_p * eps_noise).sum(axis=1)
    return dict(dz=dz, dmu=dmu, dsigma=dsigma, datt=datt, loss=fw["loss"])


def _softmax(x, axis):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def eprl_eval_forward(z, mu, sigma, eps_noise, alpha, mlp_w, mlp_b, k=100, threshold=0.5):
    """Eval branch, code/fusion_net.py:152-218 (dropout is identity in eval).

    mlp_w [C,T], mlp_b [C] are ``mlp_2d.1`` (T == 144) or ``mlp_3d.1`` (otherwise).
    Returns dict(att, labels, loss, entropy, pos_val, neg_val).  When the number of
    confident rows is neither B nor 1 the reference's mask indexing (:191) is a
    shape error; that is reported as IndexError here as well.
    """
    b = z.shape[0]
    z_p = eprl_sample_proxies(mu, sigma, eps_noise)
    z_n = _normalize_dim1(z)
    att = eprl_scores(z, mu, sigma, eps_noise)
    att_mean = att.mean(axis=2)                                # :162
    z_mean = z_n.mean(axis=2)                                  # :163  [B,T]
    pl_att = _softmax(att_mean, 1)                             # :166
    pl_feat = _softmax(z_mean, 1)                              # :167
    h = np.maximum(pl_feat, 0) @ mlp_w.T + mlp_b               # :168-171 ReLU, Linear, (Dropout), ReLU
    pl_feat = np.maximum(h, 0)
    comb = alpha * pl_att + (1 - alpha) * pl_feat              # :173
    conf = comb.max(axis=1)
    labels = comb.argmax(axis=1)                               # :177
    mask = conf > threshold                                    # :178
    if mask.sum() == 0:
        mask[conf.argmax()] = True                             # :181-182
    filt = labels[mask]
    if len(filt) not in (1, b):
        raise IndexError("shape mismatch: confident rows do not broadcast to the batch (fusion_net.py:191)")
    row_label = np.broadcast_to(filt, (b,)).astype(np.int64)
    pos, neg = eprl_split(att, row_label)
    tp, _ = topk_rows(pos, k)
    tn, _ = topk_rows(neg, k)
    loss = eprl_proxy_loss(tp, tn)
    p = _softmax(comb, 1)                                      # :127-131
    logp = np.log(p)
    entropy = float(np.mean(-np.sum(p * logp, axis=1)))
    return dict(att=att, labels=row_label, loss=loss, entropy=entropy, pos_val=tp, neg_val=tn,
                combined=comb)


# --------------------------------------------------------------------------
# North-star extension: select + gather (no reference code; oracle = topk o gather)
# --------------------------------------------------------------------------
def gather_rows(features, idx):
    """out[b, j, :] = features[b, idx[b, j], :]  (torch.gather along dim 1)."""
    b = features.shape[0]
    return features[np.arange(b)[:, None], idx.astype(np.int64), :]


def select_gather(features, scores, k):
    vals, idx = topk_rows(scores, k)
    return gather_rows(features, idx), vals, idx


# --------------------------------------------------------------------------
# Next-row: DILR Barlow-Twins cross-correlation loss
# --------------------------------------------------------------------------
def dilr_bt_loss_cross(z1, z2, common_dim, batch_size, eps=1e-5, grad_w=None):
    """code/fusion_net.py:656-677 with train-mode ``BatchNorm1d(affine=False)`` (batch mean, biased variance, eps).

    c = bn1(z1).T @ bn2(z2) / (4 batch_size); over the common block c[:dc, :dc] and the unique block c[dc:, dc:]:
    on-diagonal sum of (c_ii - 1)^2 (common) or c_ii^2 (unique), off-diagonal sum of c_ij^2, loss = on + 0.0051 off.
    Returns the six values (loss_c, on_c, off_c, loss_u, on_u, off_u); with ``grad_w`` (six weights) also the gradients
    of sum_k grad_w[k] out[k] w.r.t. z1 and z2 (closed form of what autograd does to the reference).
    """
    z1 = np.asarray(z1, dtype=np.float64)
    z2 = np.asarray(z2, dtype=np.float64)
    b, d = z1.shape
    dc = int(common_dim)

    def bn(z):
        mean = z.mean(axis=0)
        var = z.var(axis=0)                      # biased: what normalises in train mode
        inv = 1.0 / np.sqrt(var + eps)
        return (z - mean) * inv, inv

    h1, inv1 = bn(z1)
    h2, inv2 = bn(z2)
    scale = 1.0 / (batch_size * 4)
    c = (h1.T @ h2) * scale                                   # :658-661
    cc, cu = c[:dc, :dc], c[dc:, dc:]
    on_c = ((np.diag(cc) - 1.0) ** 2).sum()                   # :668
    off_c = (cc ** 2).sum() - (np.diag(cc) ** 2).sum()        # :669
    on_u = (np.diag(cu) ** 2).sum()                           # :671
    off_u = (cu ** 2).sum() - (np.diag(cu) ** 2).sum()        # :672
    out = np.array([on_c + 0.0051 * off_c, on_c, off_c, on_u + 0.0051 * off_u, on_u, off_u])
    if grad_w is None:
        return out
    gw = np.asarray(grad_w, dtype=np.float64)
    g_on_c, g_off_c = gw[0] + gw[1], 0.0051 * gw[0] + gw[2]
    g_on_u, g_off_u = gw[3] + gw[4], 0.0051 * gw[3] + gw[5]
    dcm = np.zeros_like(c)                                    # d total / d c
    dcm[:dc, :dc] = 2.0 * g_off_c * cc
    dcm[dc:, dc:] = 2.0 * g_off_u * cu
    ii = np.arange(dc)
    dcm[ii, ii] = 2.0 * g_on_c * (np.diag(cc) - 1.0)
    jj = np.arange(dc, d)
    dcm[jj, jj] = 2.0 * g_on_u * np.diag(cu)
    dcm *= scale
    dh1 = h2 @ dcm.T                                          # [b, d]
    dh2 = h1 @ dcm

    def bn_bwd(dh, h, inv):
        return inv * (dh - dh.mean(axis=0) - h * (dh * h).mean(axis=0))

    return out, bn_bwd(dh1, h1, inv1), bn_bwd(dh2, h2, inv2)
