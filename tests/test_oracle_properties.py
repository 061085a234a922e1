"""Property tests of the numpy oracle: the invariants the GPU tests rely on at full size."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import edrl_oracle as O


def _xy(seed, ns, nt, d):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((ns, d)), rng.standard_normal((nt, d)) * 1.3 + 0.2


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), ns=st.integers(2, 24), nt=st.integers(2, 24), d=st.integers(1, 40),
       scale=st.floats(0.1, 50.0))
def test_mmd_invariants(seed, ns, nt, d, scale):
    x, y = _xy(seed, ns, nt, d)
    l, m, dx, dy = O.mk_mmd_grad(x, y)
    # symmetric in its arguments
    l2, _, dy2, dx2 = O.mk_mmd_grad(y, x)
    assert np.isclose(l, l2, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dx, dx2, rtol=1e-7, atol=1e-13)
    # permutation invariant within a set (gradients permute along)
    perm = np.random.default_rng(seed + 1).permutation(ns)
    l3, _, dx3, _ = O.mk_mmd_grad(x[perm], y)
    assert np.isclose(l, l3, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dx3, dx[perm], rtol=1e-7, atol=1e-13)
    # scale covariant: the bandwidth is data derived
    l4, _, dx4, _ = O.mk_mmd_grad(scale * x, scale * y)
    assert np.isclose(l, l4, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(dx4 * scale, dx, rtol=1e-6, atol=1e-12)
    # translation invariant: gradients sum to zero, a common shift changes nothing
    assert np.abs(dx.sum(0) + dy.sum(0)).max() <= 1e-10 * max(1.0, np.abs(dx).max())
    l5 = O.mk_mmd(x + 3.0, y + 3.0)
    assert np.isclose(l, l5, rtol=1e-8, atol=1e-12)
    # the closed-form bandwidth of the centred data equals the reference's statistic
    z = np.concatenate([x, y])
    n = ns + nt
    zc = z - z.mean(0)
    assert np.isclose(O.mmd_bandwidth(z, 2.0, 5), 2 * n * (zc * zc).sum() / (n * n - n) / 4, rtol=1e-9)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), r=st.integers(1, 6), w=st.integers(1, 300), data=st.data())
def test_topk_invariants(seed, r, w, data):
    k = data.draw(st.integers(1, w))
    x = np.random.default_rng(seed).integers(-20, 20, size=(r, w)).astype(np.float32)   # plenty of ties
    v, i = O.topk_rows(x, k)
    assert np.all(np.diff(v, axis=1) <= 0)                                   # sorted descending
    np.testing.assert_array_equal(np.take_along_axis(x, i.astype(np.int64), 1), v)
    for row in range(r):
        assert len(set(i[row].tolist())) == k                                 # distinct indices
        kth = v[row, -1]
        assert (x[row] > kth).sum() <= k - 1 or (x[row] > kth).sum() < k      # nothing better was left out
        # ties: among equal values lower indices come first
        same = np.where(np.diff(v[row]) == 0)[0]
        assert np.all(i[row][same] < i[row][same + 1])


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), b=st.integers(1, 5), t=st.integers(1, 12), f=st.integers(1, 9))
def test_score_hoisting_is_exact(seed, b, t, f):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((b, t, f))
    mu, sigma = rng.standard_normal((2, f)), np.abs(rng.standard_normal((2, f))) + 0.1
    eps = rng.standard_normal((2, 7, f))
    np.testing.assert_allclose(O.eprl_scores(z, mu, sigma, eps), O.eprl_scores_hoisted(z, mu, sigma, eps),
                               rtol=1e-10, atol=1e-13)
