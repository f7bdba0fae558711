import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pf


def test_uniform_weights_identity():
    n = 1000
    w = np.full(n, 1 / n, dtype=np.float32)
    for u in (0.0, 0.3, 0.999999):
        assert np.array_equal(pf.systematic_resample(w, u), np.arange(n))


def test_one_hot_weight():
    w = np.zeros(257, dtype=np.float32); w[93] = 1
    assert (pf.systematic_resample(w, 0.5) == 93).all()


def test_tiny_tail_weights_and_clamp():
    w = np.full(64, 1e-30, dtype=np.float32); w[-1] = 1.0
    idx = pf.systematic_resample(w, 0.999999999)
    assert idx.max() == 63 and (idx == 63).all()


@settings(max_examples=60, deadline=None)
@given(st.integers(2, 700), st.integers(0, 2 ** 31), st.floats(0, 1, exclude_max=True))
def test_counts_within_one_and_sorted(n, seed, u):
    rng = np.random.default_rng(seed)
    w = rng.random(n).astype(np.float32) ** 4
    w = (w / w.sum()).astype(np.float32)
    idx = pf.systematic_resample(w, u)
    assert (np.diff(idx) >= 0).all() and idx.min() >= 0 and idx.max() < n
    q = pf.snap_weights(w)
    expect = n * q / q.sum()
    counts = np.bincount(idx, minlength=n)
    assert (np.abs(counts - expect) < 1 + 1e-6).all()


def test_snapped_cdf_is_order_independent():
    rng = np.random.default_rng(0)
    w = rng.random(32768).astype(np.float32) ** 8
    w = (w / w.sum()).astype(np.float32)
    q = pf.snap_weights(w)
    seq = np.cumsum(q)
    blocks = np.concatenate([np.cumsum(b) + o for b, o in
                             zip(np.split(q, 8), np.concatenate([[0], np.cumsum(
                                 [b.sum() for b in np.split(q, 8)])[:-1]]))])
    assert np.array_equal(seq, blocks)
    assert float(np.sort(q).sum()) == float(q.sum()) == float(q[::-1].sum()) == seq[-1]


def test_normalise_sums_to_one_and_ess():
    rng = np.random.default_rng(1)
    lw = (rng.standard_normal(4096) * 3).astype(np.float32)
    r = pf.normalise(lw)
    assert abs(float(r["w"].astype(np.float64).sum()) - 1) < 1e-5
    assert 1 <= r["ess"] <= 4096
    r2 = pf.normalise(lw + np.float32(50))
    assert np.allclose(r2["w"], r["w"], rtol=2e-5)
    assert abs(pf.normalise(np.zeros(10, np.float32))["ess"] - 10) < 1e-4


def test_weigh_cosine_and_estimate():
    rng = np.random.default_rng(2)
    emb = rng.standard_normal((16, 32)).astype(np.float32)
    tmpl = emb[5].copy()
    st_ = rng.standard_normal((6, 16)).astype(np.float32)
    r = pf.weigh(emb, tmpl, 0.1, st_)
    assert abs(r["cos"][5] - 1) < 1e-6 and r["w"].argmax() == 5
    assert r["cos"].max() <= 1 + 1e-6
    want = (r["w"][None].astype(np.float64) * st_[[0, 1, 4]]).sum(1) / r["w"].astype(np.float64).sum()
    assert np.allclose(r["est"], want, atol=1e-6)
    z = pf.weigh(np.zeros((4, 8), np.float32), tmpl[:8], 0.1)      # zero embedding: eps guard
    assert np.isfinite(z["w"]).all()


def test_propagate_and_gather():
    st_ = np.zeros((6, 8), dtype=np.float32); st_[2] = 1.0; st_[3] = -2.0
    noise = np.zeros((6, 8), dtype=np.float32)
    out = pf.propagate(st_, 1, 1, 1, 1, 0, noise=noise)
    assert (out[0] == 1).all() and (out[1] == -2).all() and (out[2] == 1).all()
    a = pf.propagate(st_, 2.0, 0.5, 0.01, 3, seed=7)
    b = pf.propagate(st_[:, 4:], 2.0, 0.5, 0.01, 3, seed=7, id_offset=4)
    assert np.array_equal(a[:, 4:], b)                              # shard == slice of the whole
    idx = np.array([7, 7, 0, 3, 3, 3, 1, 2], dtype=np.int32)
    assert np.array_equal(pf.gather_state(a, idx), a[:, idx])


def test_init_particles_centre():
    p = pf.init_particles(20000, (10, 20, 40, 60), 2.0, seed=0)
    assert abs(p[0].mean() - 30) < 0.1 and abs(p[1].mean() - 50) < 0.1
    assert abs(p[0].std() - 2) < 0.1 and (p[2:] == 0).all()
