"""GPU parity: particle-filter kernels (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import pf
from oracle.philox import philox4x32_10, philox_normals, philox_uniform53
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def test_philox_words_bit_exact():
    from vitparticlefiltertracker_b200 import ops
    cuda()
    for (off, c1, c2, c3, seed) in [(0, 0, 0, 0, 0), (1000, 7, 3, 1, 0x1234567890ABCDEF),
                                    (0xFFFFFF00, 0xFFFFFFFF, 2, 2, 2 ** 64 - 1)]:
        n = 4099
        got = ops.philox_fill(n, off, c1, c2, c3, seed).cpu().numpy().view(np.uint32)
        ids = (np.arange(n, dtype=np.uint64) + off).astype(np.uint32)
        want = philox4x32_10((ids, c1, c2, c3), (seed & 0xFFFFFFFF, seed >> 32))
        for j in range(4):
            assert np.array_equal(got[j], want[j])
    # Random123 known answers straight from the device
    got = ops.philox_fill(1, 0, 0, 0, 0, 0).cpu().numpy().view(np.uint32)[:, 0]
    assert [int(v) for v in got] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    got = ops.philox_fill(1, 0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344,
                          0x299F31D0A4093822).cpu().numpy().view(np.uint32)[:, 0]
    assert [int(v) for v in got] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normals_and_uniform():
    from vitparticlefiltertracker_b200 import ops
    cuda()
    got = ops.normals_fill(20000, 123, 9, 2, 77).cpu().numpy()
    want = philox_normals(np.arange(123, 20123), 9, 2, 77)
    # D14: fp64 transcendentals rounded once -> the normals are the oracle's, bit for bit
    assert np.array_equal(got, want), f"{(got != want).sum()} of {got.size} normals differ"
    for t in (1, 2, 99):
        assert ops.resample_uniform(5, t, 3) == philox_uniform53(t, 3, 5)


def test_init_and_propagate():
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    n = 3000
    st = torch.zeros((6, n), device=dev)
    ops.init_particles(st, 100.5, 60.25, 4.8, seed=11, stream_id=1, id_offset=64)
    want = pf.init_particles(n, (100.5 - 24, 60.25 - 12, 48, 24), 4.8, 11, 1, 64)
    assert np.array_equal(st.cpu().numpy(), want)                 # a1: bit-exact particle cloud
    rng = np.random.default_rng(0)
    st0 = rng.standard_normal((6, n)).astype(np.float32) * 50
    st = torch.from_numpy(st0).to(dev)
    ops.propagate(st, 2.4, 0.96, 0.01, seed=11, frame_idx=5, stream_id=1, id_offset=64)
    want = pf.propagate(st0, 2.4, 0.96, 0.01, 5, 11, 1, 64)
    assert np.array_equal(st.cpu().numpy(), want)                 # a3: bit-exact motion model
    assert np.array_equal(st.cpu().numpy()[5], st0[5])


def _gpu_weigh(emb, tmpl, tau, state, R=1):
    """Runs loglik -> sumexp -> weights with the particle set split over R emulated ranks."""
    from vitparticlefiltertracker_b200 import ops
    dev = emb.device
    n = emb.shape[0]
    nl = n // R
    logw = torch.empty(n, device=dev)
    mx = torch.empty(R, device=dev)
    sm = torch.empty(R, dtype=torch.float64, device=dev)
    w = torch.empty(n, device=dev)
    cdf = torch.empty(n, dtype=torch.float64, device=dev)
    tot = torch.empty(R, dtype=torch.float64, device=dev)
    stats = torch.zeros(5, dtype=torch.float64, device=dev)
    for r in range(R):
        sl = slice(r * nl, (r + 1) * nl)
        ops.loglik(emb[sl], tmpl, tau, logw[sl], mx[r:r + 1])
    for r in range(R):
        sl = slice(r * nl, (r + 1) * nl)
        ops.sumexp(logw[sl], mx, sm[r:r + 1])
    for r in range(R):
        sl = slice(r * nl, (r + 1) * nl)
        ops.weights(logw[sl], mx, sm, state[:, sl].contiguous(), w[sl], cdf[sl], tot[r:r + 1], stats)
    return dict(logw=logw, w=w, cdf=cdf, totals=tot, stats=stats, max=mx, sums=sm)


def _gpu_resample(res, state, u, R=1):
    from vitparticlefiltertracker_b200 import ops
    dev = state.device
    n = state.shape[1]
    nl = n // R
    state_all = state.view(6, R, nl).permute(1, 0, 2).contiguous()
    out = torch.empty((6, n), device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    for r in range(R):
        o = torch.empty((6, nl), device=dev)
        ops.resample(res["cdf"], res["totals"], nl, r * nl, u, state_all, o, idx[r * nl:(r + 1) * nl])
        out[:, r * nl:(r + 1) * nl] = o
    return idx, out


@pytest.mark.parametrize("n,D", [(256, 192), (4096, 768), (1000, 1024)])
def test_weigh_matches_oracle(n, D):
    dev = cuda()
    rng = np.random.default_rng(n)
    tmpl = rng.standard_normal(D).astype(np.float32)
    emb = (tmpl[None] + 0.7 * rng.standard_normal((n, D))).astype(np.float32)
    state = rng.standard_normal((6, n)).astype(np.float32) * 30
    want = pf.weigh(emb, tmpl, 0.1, state)
    got = _gpu_weigh(torch.from_numpy(emb).to(dev), torch.from_numpy(tmpl).to(dev), 0.1,
                     torch.from_numpy(state).to(dev))
    logw = got["logw"].cpu().numpy()
    # D14: fp64-accumulated cosine rounded once, fp64 exponentials rounded once -> given the
    # embeddings, log-weights and weights are the oracle's bit for bit (north_star's fp32 bar is
    # rel 1e-5)
    assert np.array_equal(logw, want["logw"]), f"{(logw != want['logw']).sum()} log-weights differ"
    w = got["w"].cpu().numpy()
    assert np.array_equal(w, want["w"]), f"{(w != want['w']).sum()} weights differ"
    s = got["stats"].cpu().numpy()
    assert abs(s[0] - 1) < 1e-5
    assert abs(1 / s[1] - want["ess"]) / want["ess"] < 1e-12
    est = np.array([s[2], s[3], s[4]]) / s[0]
    assert np.abs(est - want["est"].astype(np.float64)).max() < 1e-5   # oracle rounds est to fp32
    # the CDF is the exact scan of the snapped GPU weights
    assert np.array_equal(got["cdf"].cpu().numpy(), np.cumsum(pf.snap_weights(w)))


def _logw_cases(n, rng):
    yield "uniform", np.zeros(n, np.float32)
    one = np.full(n, -200.0, np.float32); one[n // 3] = 0.0
    yield "one_hot", one
    tail = np.full(n, -69.0, np.float32); tail[-1] = 0.0       # ~1e-30 tails
    yield "tiny_tails", tail
    yield "ties", np.repeat(rng.standard_normal(n // 8).astype(np.float32), 8)
    yield "random", (rng.standard_normal(n) * 4).astype(np.float32)
    yield "peaky", (rng.standard_normal(n) * 30).astype(np.float32)


@pytest.mark.parametrize("n", [256, 1024, 4096, 32768])
def test_resample_bit_exact_and_rank_invariant(n):
    """K4: indices are bit-exact given identical normalised weights and uniform draw, for every
    emulated rank count; the exact fp64 sums make weights themselves rank-invariant too."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    rng = np.random.default_rng(n)
    state = torch.from_numpy(rng.standard_normal((6, n)).astype(np.float32)).to(dev)
    for name, logw_np in _logw_cases(n, rng):
        for u in (0.0, 0.2169747068176081, 0.999999999999):
            ref = None
            for R in (1, 2, 8):
                nl = n // R
                logw = torch.from_numpy(logw_np).to(dev)
                mx = torch.empty(R, device=dev); sm = torch.empty(R, dtype=torch.float64, device=dev)
                w = torch.empty(n, device=dev); cdf = torch.empty(n, dtype=torch.float64, device=dev)
                tot = torch.empty(R, dtype=torch.float64, device=dev)
                stats = torch.zeros(5, dtype=torch.float64, device=dev)
                for r in range(R):
                    mx[r] = logw[r * nl:(r + 1) * nl].max()
                for r in range(R):
                    ops.sumexp(logw[r * nl:(r + 1) * nl], mx, sm[r:r + 1])
                for r in range(R):
                    sl = slice(r * nl, (r + 1) * nl)
                    ops.weights(logw[sl], mx, sm, state[:, sl].contiguous(), w[sl], cdf[sl],
                                tot[r:r + 1], stats)
                res = dict(cdf=cdf, totals=tot)
                idx, out = _gpu_resample(res, state, u, R)
                w_np = w.cpu().numpy()
                want = pf.systematic_resample(w_np, u)
                assert np.array_equal(idx.cpu().numpy(), want), (name, u, R)
                assert torch.equal(out, state[:, idx.long()])
                if ref is None:
                    ref = (w_np, idx.cpu().numpy())
                else:
                    assert np.array_equal(ref[0], w_np), (name, R)       # weights rank-invariant
                    assert np.array_equal(ref[1], idx.cpu().numpy())


def test_bad_arguments_raise():
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    st = torch.zeros((6, 8), device=dev)
    cdf = torch.ones(8, dtype=torch.float64, device=dev).cumsum(0)
    tot = cdf[-1:].clone()
    idx = torch.empty(8, dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError, match="0 <= u < 1"):
        ops.resample(cdf, tot, 8, 0, 1.0, st, st.clone(), idx)


@pytest.mark.parametrize("shards", [[1024, 1024], [1500, 548], [700, 37, 1311], [256] * 8])
def test_resample_records_uneven_shards(shards):
    """Multi-GPU form: packed records read in place, uneven shards, rank-ordered estimate sums.
    Ancestor indices and gathered state equal the oracle on the concatenated weights."""
    from vitparticlefiltertracker_b200 import ops
    from vitparticlefiltertracker_b200.exchange import HEADER
    dev = cuda()
    R, n, n_max = len(shards), sum(shards), max(shards)
    rng = np.random.default_rng(n + R)
    logw_np = (rng.standard_normal(n) * 5).astype(np.float32)
    state_np = rng.standard_normal((6, n)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(shards)]).astype(np.int32)
    shard_off = torch.from_numpy(off).to(dev)
    logw = torch.from_numpy(logw_np).to(dev)
    rec = torch.zeros((R, HEADER + 4 * n_max), dtype=torch.float64, device=dev)
    mx = torch.stack([logw[off[r]:off[r + 1]].max() for r in range(R)])
    sm = torch.empty(R, dtype=torch.float64, device=dev)
    w = torch.empty(n, device=dev)
    for r in range(R):
        ops.sumexp(logw[off[r]:off[r + 1]], mx, sm[r:r + 1])
    for r in range(R):
        nr = shards[r]
        st = rec[r, HEADER + n_max:].view(torch.float32).view(6, n_max)[:, :nr]
        st.copy_(torch.from_numpy(state_np[:, off[r]:off[r + 1]].copy()))
        ops.weights(logw[off[r]:off[r + 1]], mx, sm, st, w[off[r]:off[r + 1]],
                    rec[r, HEADER:HEADER + nr], rec[r, 0:1], rec[r, 1:6], accumulate=False)
    want = pf.normalise(logw_np)
    assert np.array_equal(w.cpu().numpy(), want["w"])
    for u in (0.0, 0.6180339887, 0.999999999999):
        idx_all = pf.systematic_resample(want["w"], u)
        for r in range(R):
            nr = shards[r]
            out = torch.full((6, nr + 3), -1.0, device=dev)[:, :nr]      # a strided destination
            idx = torch.empty(nr, dtype=torch.int32, device=dev)
            stats = torch.empty(5, dtype=torch.float64, device=dev)
            ops.resample_records(rec, n_max, shard_off, nr, int(off[r]), u, out, idx, stats_out=stats)
            assert np.array_equal(idx.cpu().numpy(), idx_all[off[r]:off[r + 1]]), (r, u)
            assert np.array_equal(out.cpu().numpy(), state_np[:, idx_all[off[r]:off[r + 1]]])
            want_stats = torch.zeros(5, dtype=torch.float64, device=dev)
            for q in range(R):                               # the kernel sums in rank order
                want_stats += rec[q, 1:6]
            assert torch.equal(stats, want_stats)
