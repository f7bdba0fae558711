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


def _rank_main(rank, world, port, n_global, u, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vitparticlefiltertracker_b200.exchange import Exchange, shard_range
        rng = np.random.default_rng(7)
        logw = (rng.standard_normal(n_global) * 6).astype(np.float32)
        state = rng.standard_normal((6, n_global)).astype(np.float32)
        off, n = shard_range(n_global, rank, world)
        sl = slice(off, off + n)
        xch = Exchange(dist.group.WORLD, n, "cpu")
        # a6: local max -> all-gather -> exact local exp-sum -> all-gather -> weights + local CDF
        xch.max_local[0] = float(logw[sl].max())
        m = np.float32(xch.gather_max().numpy().max())
        e = np.exp((logw[sl] - m).astype(np.float32)).astype(np.float32)
        xch.sum_local[0] = float((np.rint(e.astype(np.float64) * 2.0 ** pf.SUM_GRID_BITS)
                                  * 2.0 ** -pf.SUM_GRID_BITS).sum())
        ssum = float(xch.gather_sum().numpy().sum())
        lse = np.float32(np.float64(m) + np.log(ssum))
        w = np.exp((logw[sl] - lse).astype(np.float32)).astype(np.float32)
        total, cdf, st_slot = xch.local_views()
        cdf.copy_(torch.from_numpy(np.cumsum(pf.snap_weights(w))))
        total[0] = cdf[-1]
        st_slot.copy_(torch.from_numpy(state[:, sl].copy()))
        totals, cdf_all, state_all = xch.gather_records()
        # a7/a8 exactly as resample_kernel does it
        offs = np.concatenate([[0.0], np.cumsum(totals.numpy())[:-1]])
        C = (cdf_all.numpy() + offs[:, None]).reshape(-1)
        step = C[-1] / np.float64(n_global)
        p = (np.arange(off, off + n, dtype=np.float64) + np.float64(u)) * step
        a = np.minimum(np.searchsorted(C, p, side="right"), n_global - 1)
        gathered = state_all.numpy().transpose(1, 0, 2).reshape(6, -1)[:, a]
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=w, idx=a, state=gathered, C=C)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [64, 4096])
def test_two_rank_protocol_matches_single_rank_oracle(tmp_path, n_global):
    world, u = 2, 0.2169747068176081
    mp.spawn(_rank_main, args=(world, _free_port(), n_global, u, str(tmp_path)), nprocs=world,
             join=True)
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


def test_shard_range_and_single_rank_aliasing():
    from vitparticlefiltertracker_b200.exchange import Exchange, shard_range
    assert shard_range(4096, 3, 8) == (1536, 512)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    x = Exchange(None, 8, "cpu")
    assert x.max_all.data_ptr() == x.max_local.data_ptr()
    t, c, s = x.local_views()
    assert t.numel() == 1 and c.numel() == 8 and s.shape == (6, 8)
    tt, cc, ss = x.gather_records()
    assert cc.shape == (1, 8) and ss.shape == (1, 6, 8)
