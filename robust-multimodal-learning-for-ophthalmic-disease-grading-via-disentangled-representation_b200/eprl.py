"""Part B host side: the Essence-Point module ``EPRL`` with the reference's constructor,
parameter names and ``forward`` contract (reference: code/fusion_net.py:63-255), its scoring /
selection running on the sm_100a kernels of ``csrc/eprl.cu``.

What stays torch: the 3-layer encoder MLP (stock GEMMs), softplus on the
proxies and the [B, C]-sized pseudo-label arithmetic of the eval branch.  What is replaced:
normalise-over-tokens + token mean (never materialising [B,C,T,S]), proxy sampling +
sample-dim normalisation, the score contraction, the label-addressed split (no masked_select,
no per-label Python loop), top-k, the proxy loss, and all of their backward passes.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

SELF_TOPK = 100  # hard-coded in the reference (code/fusion_net.py:199,236)


def _f32c(t):
    return t.to(torch.float32).contiguous()


# ----------------------------------------------------------------------------- functional ops
class _ScoreFunction(torch.autograd.Function):
    """att[b,c,s] = mean_t normalize(z, dim=1)[b,t,:] . normalize(mu + sigma*eps, dim=1)[c,s,:]
    (code/fusion_net.py:143-150, 221-225) with the token mean hoisted in front of the contraction."""

    @staticmethod
    def forward(ctx, z, mu, sigma, eps):
        lib = _lib.load()
        z, mu, sigma, eps = _f32c(z), _f32c(mu), _f32c(sigma), _f32c(eps)
        B, T, Fd = z.shape
        C, S, F2 = eps.shape
        if F2 != Fd or mu.shape != (C, Fd) or sigma.shape != (C, Fd):
            raise RuntimeError(f"EPRL score: shape mismatch z{tuple(z.shape)} mu{tuple(mu.shape)} "
                               f"sigma{tuple(sigma.shape)} eps{tuple(eps.shape)}")
        dev = z.device
        st = _lib.stream_and_device(z)
        zbar = torch.empty(B, Fd, device=dev)
        colsum = torch.empty(B, Fd, device=dev)
        colnorm = torch.empty(B, Fd, device=dev)
        z_pn = torch.empty(C, S, Fd, device=dev)
        pnorm = torch.empty(C, Fd, device=dev)
        att = torch.empty(B, C, S, device=dev)
        _lib.check(lib.edrl_token_stats_fwd(z.data_ptr(), B, T, Fd, zbar.data_ptr(), colsum.data_ptr(),
                                            colnorm.data_ptr(), st))
        _lib.check(lib.edrl_proxy_normalize_fwd(mu.data_ptr(), sigma.data_ptr(), eps.data_ptr(), C, S, Fd,
                                                z_pn.data_ptr(), pnorm.data_ptr(), st))
        _lib.check(lib.edrl_score_fwd(zbar.data_ptr(), z_pn.data_ptr(), B, C * S, Fd, att.data_ptr(), st))
        ctx.save_for_backward(z, mu, sigma, eps, zbar, colsum, colnorm, z_pn, pnorm)
        ctx.mark_non_differentiable(colnorm)
        return att, colnorm

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, datt, _dcolnorm):
        lib = _lib.load()
        z, mu, sigma, eps, zbar, colsum, colnorm, z_pn, pnorm = ctx.saved_tensors
        B, T, Fd = z.shape
        C, S, _ = eps.shape
        dev = z.device
        datt = _f32c(datt)
        st = _lib.stream_and_device(z)
        need_z = ctx.needs_input_grad[0]
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dzbar = torch.empty(B, Fd, device=dev) if need_z else None
        dz_pn = torch.empty(C, S, Fd, device=dev) if need_p else None
        _lib.check(lib.edrl_score_bwd(datt.data_ptr(), zbar.data_ptr(), z_pn.data_ptr(), B, C * S, Fd,
                                      _lib.ptr(dzbar), _lib.ptr(dz_pn), st))
        dz = dmu = dsigma = None
        if need_z:
            dz = torch.empty_like(z)
            _lib.check(lib.edrl_token_stats_bwd(z.data_ptr(), colsum.data_ptr(), colnorm.data_ptr(),
                                                dzbar.data_ptr(), B, T, Fd, dz.data_ptr(), st))
        if need_p:
            dmu = torch.empty_like(mu)
            dsigma = torch.empty_like(sigma)
            _lib.check(lib.edrl_proxy_normalize_bwd(mu.data_ptr(), sigma.data_ptr(), eps.data_ptr(),
                                                    pnorm.data_ptr(), dz_pn.data_ptr(), C, S, Fd, dmu.data_ptr(),
                                                    dsigma.data_ptr(), st))
        return dz, dmu, dsigma, None


class _SelectLossFunction(torch.autograd.Function):
    """proxy_loss = mean_b exp(-mean(top_k att[b, y_b, :]) + mean(top_k concat_{c != y_b} att[b, c, :]))
    (code/fusion_net.py:227-243).  Also returns the selected values / indices."""

    @staticmethod
    def forward(ctx, att, y, k, sorted_):
        lib = _lib.load()
        att = _f32c(att)
        B, C, S = att.shape
        y = y.to(device=att.device, dtype=torch.int64).contiguous()
        dev = att.device
        st = _lib.stream_and_device(att)
        vals = torch.empty(2, B, k, device=dev)
        idx = torch.empty(2, B, k, dtype=torch.int32, device=dev)
        loss = torch.empty((), device=dev)
        rowexp = torch.empty(B, device=dev)
        _lib.check(lib.edrl_select_topk_fwd(att.data_ptr(), y.data_ptr(), B, C, S, k, int(bool(sorted_)),
                                            vals[0].data_ptr(),
                                            idx[0].data_ptr(), vals[1].data_ptr(), idx[1].data_ptr(), st))
        _lib.check(lib.edrl_proxy_loss_fwd(vals[0].data_ptr(), vals[1].data_ptr(), B, k, loss.data_ptr(),
                                           rowexp.data_ptr(), st))
        ctx.save_for_backward(rowexp, idx, y)
        ctx.shape = (B, C, S, k)
        ctx.mark_non_differentiable(vals, idx)
        return loss, vals, idx

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss, _gv, _gi):
        lib = _lib.load()
        rowexp, idx, y = ctx.saved_tensors
        B, C, S, k = ctx.shape
        g = _f32c(gloss)
        st = _lib.stream_and_device(rowexp)
        datt = torch.empty(B, C, S, device=rowexp.device)
        _lib.check(lib.edrl_select_loss_bwd(rowexp.data_ptr(), idx[0].data_ptr(), idx[1].data_ptr(), y.data_ptr(),
                                            g.data_ptr(), B, C, S, k, datt.data_ptr(), st))
        return datt, None, None, None


class _EssenceTrainFunction(torch.autograd.Function):
    """The whole train-branch loss after the encoder (code/fusion_net.py:137-150, 220-243) as one C-ABI call forward
    and one backward: proxy_loss(z, proxies, eps, y).  Same kernels as _ScoreFunction + _SelectLossFunction; softplus
    on the raw proxy half is applied inside the proxy kernel."""

    @staticmethod
    def forward(ctx, z, proxies, eps, y, k):
        lib = _lib.load()
        z, proxies, eps = _f32c(z), _f32c(proxies), _f32c(eps)
        B, T, Fd = z.shape
        C, S, F2 = eps.shape
        if F2 != Fd or proxies.shape != (C, 2 * Fd):
            raise RuntimeError(f"EPRL: shape mismatch z{tuple(z.shape)} proxies{tuple(proxies.shape)} "
                               f"eps{tuple(eps.shape)}")
        y = y.to(device=z.device, dtype=torch.int64).contiguous()
        st = _lib.stream_and_device(z)
        saved = torch.empty(int(lib.edrl_essence_saved_floats(B, T, Fd, C, S, k)), device=z.device)
        loss = torch.empty((), device=z.device)
        _lib.check(lib.edrl_essence_train_fwd(z.data_ptr(), proxies.data_ptr(), eps.data_ptr(), y.data_ptr(), B, T, Fd,
                                              C, S, k, loss.data_ptr(), saved.data_ptr(), st))
        ctx.save_for_backward(z, proxies, eps, y, saved)
        ctx.dims = (B, T, Fd, C, S, k)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss):
        lib = _lib.load()
        z, proxies, eps, y, saved = ctx.saved_tensors
        B, T, Fd, C, S, k = ctx.dims
        g = _f32c(gloss)
        st = _lib.stream_and_device(z)
        scratch = torch.empty(int(lib.edrl_essence_scratch_floats(B, T, Fd, C, S, k)), device=z.device)
        dz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        dprox = torch.empty_like(proxies) if ctx.needs_input_grad[1] else None
        _lib.check(lib.edrl_essence_train_bwd(z.data_ptr(), proxies.data_ptr(), eps.data_ptr(), y.data_ptr(), B, T, Fd,
                                              C, S, k, saved.data_ptr(), g.data_ptr(), scratch.data_ptr(),
                                              _lib.ptr(dz), _lib.ptr(dprox), st))
        return dz, dprox, None, None, None


def essence_train_loss(z, proxies, eps, y, k=SELF_TOPK):
    """proxy_loss of the train branch from the encoder output z [B,T,F], the raw proxies parameter [C,2F]
    (mu | pre-softplus sigma), the noise eps [C,S,F] and labels y [B] -- one fused call each way."""
    _lib.require_cuda(z, proxies, eps)
    if z.dim() != 3:
        raise RuntimeError(f"EPRL expects token features [B, T, F], got {tuple(z.shape)}")
    if k > eps.shape[1]:
        raise RuntimeError("selected index k out of range")
    return _EssenceTrainFunction.apply(z, proxies, eps, y, int(k))


def essence_scores(z, mu, sigma, eps):
    """Differentiable att [B,C,S] (see _ScoreFunction); also returns the per-(b,f) token norms."""
    _lib.require_cuda(z, mu, sigma, eps)
    if z.dim() != 3:
        raise RuntimeError(f"EPRL expects token features [B, T, F], got {tuple(z.shape)}")
    return _ScoreFunction.apply(z, mu, sigma, eps)


def essence_select_loss(att, y, k=SELF_TOPK, sorted=True):
    """(proxy_loss, top values [2,B,k] (pos, neg), top indices [2,B,k]).  ``sorted=False`` returns the same
    selection in ascending index order (the loss only averages it)."""
    _lib.require_cuda(att)
    B, C, S = att.shape
    if k > S:
        # torch.topk(att_positive, 100, dim=1) in the reference raises this RuntimeError
        raise RuntimeError("selected index k out of range")
    return _SelectLossFunction.apply(att, y, int(k), bool(sorted))


def topk_rows(x, k, sorted=True):
    """torch.topk(x, k, dim=1) on the select kernel: (values [R,k] descending, indices int32 [R,k]),
    ties lowest-index-first; ``sorted=False`` gives the same set in ascending index order.
    Not differentiable (use select_gather for the gather path)."""
    _lib.require_cuda(x)
    if x.dim() != 2:
        raise RuntimeError("topk_rows expects a 2-D tensor")
    lib = _lib.load()
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(1) != 1:
        x = x.contiguous()
    R, W = x.shape
    if k > W or k < 1:
        raise RuntimeError("selected index k out of range")
    st = _lib.stream_and_device(x)
    vals = torch.empty(R, k, device=x.device)
    idx = torch.empty(R, k, dtype=torch.int32, device=x.device)
    if R:
        _lib.check(lib.edrl_topk_rows(x.data_ptr(), R, W, x.stride(0), k, int(bool(sorted)), vals.data_ptr(),
                                      idx.data_ptr(), st), RuntimeError)
    return vals, idx


class _GatherRowsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        lib = _lib.load()
        feat = _f32c(features)
        idx = idx.to(torch.int32).contiguous()
        B, T, D = feat.shape
        k = idx.shape[1]
        st = _lib.stream_and_device(feat)
        out = torch.empty(B, k, D, device=feat.device)
        _lib.check(lib.edrl_gather_rows_fwd(feat.data_ptr(), idx.data_ptr(), B, T, D, k, out.data_ptr(), st))
        ctx.save_for_backward(idx)
        ctx.shape = (B, T, D, k)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        (idx,) = ctx.saved_tensors
        B, T, D, k = ctx.shape
        dout = _f32c(dout)
        st = _lib.stream_and_device(dout)
        dfeat = torch.empty(B, T, D, device=dout.device)
        _lib.check(lib.edrl_gather_rows_bwd(dout.data_ptr(), idx.data_ptr(), B, T, D, k, dfeat.data_ptr(), st))
        return dfeat, None


def gather_rows(features, idx):
    """out[b, j, :] = features[b, idx[b, j], :] (differentiable w.r.t. features; idx rows distinct)."""
    _lib.require_cuda(features, idx)
    return _GatherRowsFunction.apply(features, idx)


def select_gather(features, scores, k, sorted=True):
    """North-star extension (no reference code): top-k over scores [B,T], gather features [B,T,D].
    Returns (gathered [B,k,D], values [B,k], indices int32 [B,k])."""
    vals, idx = topk_rows(scores, k, sorted)
    return gather_rows(features, idx), vals, idx


# ----------------------------------------------------------------------------- the module
class EPRL(nn.Module):
    """Essence-Point Representation Learning -- same constructor, parameters, state_dict keys and
    forward contract as the reference class (code/fusion_net.py:63-255).

    ``noise="reference"`` draws the proxy noise exactly like the reference (CPU default generator,
    copied to the device; eval mode reseeds the *global* generator with ``seed``);
    ``noise="device"`` draws it on the GPU (different stream, no H2D copy).
    """

    def __init__(self, x_dim, z_dim=256, beta=1e-2, sample_num=50, topk=1, num_classes=3, seed=1, batch_size=16,
                 noise="reference", validate_labels="lazy"):
        super().__init__()
        self.beta = beta
        self.sample_num = sample_num
        self.topk = 1                      # the reference ignores its `topk` argument (code/fusion_net.py:77)
        self.num_classes = num_classes
        self.seed = seed
        self.z_dim = z_dim
        self.encoder = nn.Sequential(
            nn.Linear(x_dim, z_dim * 2), nn.ReLU(inplace=True), nn.Dropout(0.2),
            nn.Linear(z_dim * 2, z_dim * 2), nn.ReLU(inplace=True), nn.Dropout(0.2),
            nn.Linear(z_dim * 2, z_dim),
        )
        self.batch_size = batch_size
        self.decoder_logits = nn.Linear(z_dim, num_classes)
        self.mlp_2d = nn.Sequential(nn.ReLU(), nn.Linear(144, num_classes), nn.Dropout(0.2), nn.ReLU())
        self.mlp_3d = nn.Sequential(nn.ReLU(), nn.Linear(216, num_classes), nn.Dropout(0.2), nn.ReLU())
        self.proxies = nn.Parameter(torch.empty([num_classes, z_dim * 2]))
        torch.nn.init.xavier_uniform_(self.proxies, gain=1.0)
        self.proxies_dict = {"0": 0, "1": 1}
        self.alpha = nn.Parameter(torch.tensor(0.5))
        if noise not in ("reference", "device"):
            raise ValueError("noise must be 'reference' or 'device'")
        self.noise = noise
        if validate_labels not in (True, False, "lazy"):
            raise ValueError("validate_labels must be True, False or 'lazy'")
        self.validate_labels = validate_labels
        self._verdict_host = None
        self._verdict_event = None
        self._verdict_batch = 0
        self.self_topk = SELF_TOPK
        self._eval_gen = None

    # ---- small public helpers of the reference type -------------------------------------------
    def gaussian_noise(self, samples, K, seed):
        """code/fusion_net.py:105-110."""
        dev = self.proxies.device
        if self.noise == "device":
            if self.training:
                return torch.randn(*samples, K, device=dev)
            if self._eval_gen is None or self._eval_gen.device != dev:
                self._eval_gen = torch.Generator(device=dev)
            self._eval_gen.manual_seed(seed)
            return torch.randn(*samples, K, device=dev, generator=self._eval_gen)
        if self.training:
            return torch.normal(torch.zeros(*samples, K), torch.ones(*samples, K)).to(dev, non_blocking=True)
        return torch.normal(torch.zeros(*samples, K), torch.ones(*samples, K),
                            generator=torch.manual_seed(seed)).to(dev, non_blocking=True)

    def encoder_result(self, x):
        return self.encoder(x)

    def encoder_proxies(self):
        mu_proxy = self.proxies[:, :self.z_dim]
        sigma_proxy = F.softplus(self.proxies[:, self.z_dim:])
        return mu_proxy, sigma_proxy

    def estimate_v(self, z_proxy, epsilon=1e-8):
        var = torch.var(z_proxy, dim=1, unbiased=False)
        return torch.clamp(2 * var / (var - 1 + epsilon), min=2)

    def entropy_regularization(self, logits):
        p = torch.softmax(logits, dim=1)
        log_p = torch.log_softmax(logits, dim=1)
        return (-torch.sum(p * log_p, dim=1)).mean()

    # ---- label handling ------------------------------------------------------------------------
    def _proxy_indices(self, labels):
        """proxies_dict lookup (code/fusion_net.py:101,186,227): labels outside {0,1} raise KeyError.

        ``validate_labels="lazy"`` (default): the range check runs on the device, its verdict goes to pinned host memory
        with an asynchronous copy and is looked at by the NEXT call (or by ``check_labels()``), so the training step never
        waits for the GPU -- the reference pays one device-to-host sync per label here.  ``True``: checked at once (one
        sync per call), the reference's timing of the KeyError.  ``False``: not checked (the kernels clamp)."""
        labels = labels.to(self.proxies.device).long().reshape(-1)
        mode = self.validate_labels
        if mode and labels.numel():
            lo, hi = torch.aminmax(labels)
            if mode == "lazy":
                self.check_labels()
                self._post_verdict(torch.stack([((lo < 0) | (hi > 1)).long(), torch.where(lo < 0, lo, hi),
                                                torch.zeros_like(lo)]))
            else:
                lo, hi = torch.stack([lo, hi]).tolist()           # one sync for both
                for bad in (lo, hi):
                    if str(bad) not in self.proxies_dict:
                        raise KeyError(str(bad))
        return labels

    def _post_verdict(self, verdict):
        """verdict: int64[3] on the device = (label out of range?, the offending label, eval: unbroadcastable count or 0)"""
        if self._verdict_host is None:
            self._verdict_host = torch.zeros(3, dtype=torch.int64).pin_memory()
        self._verdict_host.copy_(verdict, non_blocking=True)
        self._verdict_event = torch.cuda.Event()
        self._verdict_event.record(torch.cuda.current_stream(verdict.device))

    def check_labels(self, wait=False):
        """Raise what the lazy checks of an earlier forward found (KeyError for a label outside proxies_dict,
        IndexError for the eval branch's unbroadcastable pseudo-label count).  ``wait=True`` blocks until that forward has
        finished on the device; otherwise a verdict that is still in flight is left for the next call."""
        ev = self._verdict_event
        if ev is None:
            return
        if wait:
            ev.synchronize()
        elif not ev.query():
            return
        self._verdict_event = None
        bad, label, count = self._verdict_host.tolist()
        if bad:
            raise KeyError(str(label))
        if count:
            raise IndexError(f"shape mismatch: indexing tensors could not be broadcast together with shapes "
                             f"[{self._verdict_batch}], [{count}]")

    # ---- forward --------------------------------------------------------------------------------
    def forward(self, x, y=None):
        _lib.require_cuda(x, self.proxies)
        z = self.encoder_result(x)
        mu_proxy, sigma_proxy = self.encoder_proxies()
        eps_proxy = self.gaussian_noise(samples=([self.num_classes, self.sample_num]), K=self.z_dim, seed=self.seed)
        B = x.shape[0]

        if not self.training:
            att, colnorm = essence_scores(z, mu_proxy, sigma_proxy, eps_proxy)
            threshold = 0.5
            att_mean = att.mean(dim=2)                                        # :162
            z_mean = self._token_featmean(z, colnorm)                         # :163  [B, T]
            pseudo_att = torch.softmax(att_mean, dim=1)
            pseudo_feat = torch.softmax(z_mean, dim=1)
            pseudo_feat = self.mlp_2d(pseudo_feat) if pseudo_feat.shape[1] == 144 else self.mlp_3d(pseudo_feat)
            combined = self.alpha * pseudo_att + (1 - self.alpha) * pseudo_feat
            confidence, labels = torch.max(combined, dim=1)
            mask = confidence > threshold
            if self.validate_labels == "lazy":
                # the reference's `.item()` (:181), boolean indexing (:184) and per-label Python loop (:186) without a
                # device-to-host sync: no confident sample -> the most confident one; the surviving pseudo-labels must
                # number 1 (broadcast to every row) or B (:191), anything else is reported by the next call
                self.check_labels()
                none = ~mask.any()
                mask = mask | (none & F.one_hot(confidence.argmax(), B).bool())
                count = mask.sum()
                first = labels[mask.to(torch.uint8).argmax()]
                row_labels = torch.where(count == B, labels, first.expand(B)).contiguous()
                sel = torch.where(mask, labels, first.expand(B))
                lo, hi = torch.aminmax(sel)
                self._verdict_batch = B
                self._post_verdict(torch.stack([((lo < 0) | (hi > 1)).long(), torch.where(lo < 0, lo, hi),
                                                torch.where((count == 1) | (count == B), torch.zeros_like(count), count)]))
            else:
                if mask.sum().item() == 0:
                    mask[confidence.argmax()] = True
                filtered = labels[mask]
                proxy_indices = self._proxy_indices(filtered)
                if proxy_indices.numel() not in (1, B):
                    # the reference's mask[arange(B), proxy_indices] broadcast fails here (:191)
                    raise IndexError(f"shape mismatch: indexing tensors could not be broadcast together with shapes "
                                     f"[{B}], [{proxy_indices.numel()}]")
                row_labels = proxy_indices.expand(B).contiguous()
            proxy_loss, _, _ = essence_select_loss(att, row_labels, self.self_topk, sorted=False)
            entropy_loss = self.entropy_regularization(combined)
            return mu_proxy.repeat(B, 1, 1), sigma_proxy.repeat(B, 1, 1), proxy_loss, z, entropy_loss

        if B != self.batch_size and self.batch_size != 1 and B != 1:
            # expand(self.batch_size, ...) against a [B,1,T,F] operand (code/fusion_net.py:221-223); B == 1 broadcasts
            # there (batch_size identical rows, the label index broadcasts at :231): the loss equals the single row's
            raise RuntimeError(f"The size of tensor a ({B}) must match the size of tensor b ({self.batch_size}) "
                               "at non-singleton dimension 0")
        if y is None:
            raise TypeError("'NoneType' object is not iterable")      # `for y_item in y` in the reference
        labels = self._proxy_indices(y)
        if labels.numel() != B:
            raise IndexError(f"shape mismatch: indexing tensors could not be broadcast together with shapes "
                             f"[{B}], [{labels.numel()}]")
        proxy_loss = essence_train_loss(z, self.proxies, eps_proxy, labels, self.self_topk)
        return mu_proxy.repeat(B, 1, 1), sigma_proxy.repeat(B, 1, 1), proxy_loss, z

    @staticmethod
    def _token_featmean(z, colnorm):
        """mean_f normalize(z, dim=1)[b,t,f] (code/fusion_net.py:163); used by the eval-only
        pseudo-label branch, not differentiated."""
        lib = _lib.load()
        zf = _f32c(z.detach())
        B, T, Fd = zf.shape
        out = torch.empty(B, T, device=zf.device)
        st = _lib.stream_and_device(zf)
        _lib.check(lib.edrl_token_featmean(zf.data_ptr(), colnorm.data_ptr(), B, T, Fd, out.data_ptr(), st))
        return out
