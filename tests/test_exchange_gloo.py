"""world_size-2 (gloo, CPU) test of the particle-sharded exchange protocol: two ranks, each
holding half of the particles, must reproduce the single-rank oracle bit for bit (weights, CDF,
ancestor indices, gathered state).  The kernels' arithmetic is stood in for by the oracle's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pf


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n_global, u, shards, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vitparticlefiltertracker_b200.exchange import Exchange, shard_range
        rng = np.random.default_rng(7)
        logw = (rng.standard_normal(n_global) * 6).astype(np.float32)
        state = rng.standard_normal((6, n_global)).astype(np.float32)
        off, n = shard_range(n_global, rank, world, shards)
        sl = slice(off, off + n)
        xch = Exchange(dist.group.WORLD, n, "cpu", shards)
        assert xch.id_offset == off
        # a6: local max -> all-gather -> exact local exp-sum -> all-gather -> weights + local CDF
        xch.max_local[0] = float(logw[sl].max())
        m = np.float32(xch.gather_max().numpy().max())
        e = pf.exp_f32((logw[sl] - m).astype(np.float32))
        xch.sum_local[0] = float((np.rint(e.astype(np.float64) * 2.0 ** pf.SUM_GRID_BITS)
                                  * 2.0 ** -pf.SUM_GRID_BITS).sum())
        ssum = float(xch.gather_sum().numpy().sum())
        lse = np.float32(np.float64(m) + np.log(ssum))
        w = pf.exp_f32((logw[sl] - lse).astype(np.float32))
        total, cdf, st_slot = xch.local_views()
        cdf.copy_(torch.from_numpy(np.cumsum(pf.snap_weights(w))))
        total[0] = cdf[-1]
        st_slot.copy_(torch.from_numpy(state[:, sl].copy()))      # the state lives in the record
        xch.stats_view().copy_(torch.tensor([w.astype(np.float64).sum(), 0, 0, 0, float(rank)]))
        rec_all = xch.gather_records()
        assert rec_all.data_ptr() == xch.rec_all.data_ptr()       # read in place, no repacking
        totals, cdfs, states, stats = xch.unpack()
        # a7/a8 exactly as resample_kernel does it (global CDF = rank prefix + local CDF)
        offs = np.concatenate([[0.0], np.cumsum(totals.numpy())[:-1]])
        C = np.concatenate([c.numpy() + o for c, o in zip(cdfs, offs)])
        step = C[-1] / np.float64(n_global)
        p = (np.arange(off, off + n, dtype=np.float64) + np.float64(u)) * step
        a = np.minimum(np.searchsorted(C, p, side="right"), n_global - 1)
        gathered = np.concatenate([s_.numpy() for s_ in states], axis=1)[:, a]
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=w, idx=a, state=gathered, C=C,
                 stats=stats.numpy().sum(0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_global,shards", [(64, None), (4096, None), (4096, [2600, 1496]),
                                             (101, [37, 64])])
def test_two_rank_protocol_matches_single_rank_oracle(tmp_path, n_global, shards):
    world, u = 2, 0.2169747068176081
    mp.spawn(_rank_main, args=(world, _free_port(), n_global, u, shards, str(tmp_path)),
             nprocs=world, join=True)
    rng = np.random.default_rng(7)
    logw = (rng.standard_normal(n_global) * 6).astype(np.float32)
    state = rng.standard_normal((6, n_global)).astype(np.float32)
    ref = pf.normalise(logw)
    idx = pf.systematic_resample(ref["w"], u)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([p["w"] for p in parts]), ref["w"])
    assert np.array_equal(parts[0]["C"], np.cumsum(pf.snap_weights(ref["w"])))
    assert np.array_equal(parts[0]["C"], parts[1]["C"])
    assert np.array_equal(np.concatenate([p["idx"] for p in parts]), idx)
    assert np.array_equal(np.concatenate([p["state"] for p in parts], axis=1), state[:, idx])
    assert np.array_equal(parts[0]["stats"], parts[1]["stats"]) and parts[0]["stats"][4] == 1.0


def test_shard_helpers_and_single_rank_aliasing():
    from vitparticlefiltertracker_b200.exchange import (Exchange, HEADER, balanced_shards,
                                                        shard_range, shard_sizes)
    assert shard_range(4096, 3, 8) == (1536, 512)
    assert shard_range(100, 1, 2, [30, 70]) == (30, 70)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    with pytest.raises(ValueError):
        shard_sizes(10, 2, [4, 5])
    b = balanced_shards(32768, [160.0, 170.0, 165.0, 158.0, 161.0, 175.0, 160.0, 162.0])
    assert sum(b) == 32768 and all(v % 8 == 0 for v in b) and b[5] == min(b) and b[3] == max(b)
    assert balanced_shards(4096, [1.0, 1.0]) == [2048, 2048]
    x = Exchange(None, 8, "cpu")
    assert x.max_all.data_ptr() == x.max_local.data_ptr()
    t, c, s = x.local_views()
    assert t.numel() == 1 and c.numel() == 8 and s.shape == (6, 8) and x.rec_len == HEADER + 32
    assert x.gather_records().data_ptr() == x.rec_local.data_ptr()
    y = Exchange(None, 5, "cpu")            # odd count: the fp32 state section stays 8-byte aligned
    assert y.state_view().shape == (6, 5) and y.state_view().stride() == (5, 1)
