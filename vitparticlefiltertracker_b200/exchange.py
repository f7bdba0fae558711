"""Host-side exchange protocol of the particle-sharded multi-GPU path.

Device-agnostic torch code: NCCL on GPUs in the product, gloo on CPU in tests
(tests/test_exchange_gloo.py, world_size 2).  Three tiny all-gathers per frame; every exchanged
quantity is exact (a max, or an fp64 sum of grid-snapped terms), so the result does not depend
on the number of ranks or on how the particles are split between them.

Per-rank record, fp64 slots (include/vpf.h, vpf_resample_records), n_max = largest shard:

    [0] total | [1..6) estimate sums | [6..8) pad | [8, 8+n) local inclusive CDF |
    [8+n_max, 8+4*n_max) state fp32 [6][n_max]

Every rank's record has the same length, so one `all_gather_into_tensor` moves them, and the
resampling kernel reads the receive buffer in place (no repacking).  The rank's particle state
lives *inside* its own record (`state_view()`), so nothing is copied before the gather either.
"""
from __future__ import annotations

import numpy as np
import torch

HEADER = 8          # VPF_RECORD_HEADER


def shard_sizes(n_global: int, world: int, shards=None):
    """Particles per rank: equal shards by default, or an explicit list (uneven shards let a slower
    GPU carry fewer particles; the filter's result does not depend on the split)."""
    if shards is None:
        if n_global % world:
            raise ValueError("n_particles must be divisible by the number of ranks "
                             "(or pass explicit shards)")
        return [n_global // world] * world
    shards = [int(s) for s in shards]
    if len(shards) != world or sum(shards) != n_global or min(shards) < 1:
        raise ValueError(f"shards {shards} must be {world} positive counts summing to {n_global}")
    return shards


def shard_range(n_global: int, rank: int, world: int, shards=None):
    """(first global particle id, count) of `rank`."""
    sz = shard_sizes(n_global, world, shards)
    return sum(sz[:rank]), sz[rank]


def balanced_shards(n_global: int, ms_per_rank, quantum: int = 8):
    """Shard sizes proportional to each rank's measured speed (1 / ms per particle), in multiples
    of `quantum`, summing to n_global."""
    speed = 1.0 / np.asarray(ms_per_rank, dtype=np.float64)
    ideal = n_global * speed / speed.sum()
    sz = np.maximum(quantum, (np.floor(ideal / quantum) * quantum).astype(np.int64))
    rem = n_global - int(sz.sum())
    order = np.argsort(-(ideal - sz))                   # largest shortfall first
    k = 0
    while rem != 0:
        step = min(quantum, abs(rem)) * (1 if rem > 0 else -1)
        i = order[k % len(sz)]
        if sz[i] + step >= 1:
            sz[i] += step
            rem -= step
        k += 1
    return [int(v) for v in sz]


class Exchange:
    def __init__(self, process_group, n_local: int, device, shards=None):
        import torch.distributed as dist
        self.dist = dist
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.rank = dist.get_rank(process_group) if process_group is not None else 0
        R = self.world
        self.shards = [int(n_local)] * R if shards is None else [int(s) for s in shards]
        if len(self.shards) != R or self.shards[self.rank] != n_local:
            raise ValueError("shards must list one count per rank, this rank's equal to n_local")
        self.n = n_local
        self.n_max = max(self.shards)
        self.rec_len = HEADER + 4 * self.n_max
        off = np.concatenate([[0], np.cumsum(self.shards)]).astype(np.int32)
        self.id_offset = int(off[self.rank])
        self.shard_off = torch.from_numpy(off).to(device)
        self.max_local = torch.empty((1,), dtype=torch.float32, device=device)
        self.sum_local = torch.empty((1,), dtype=torch.float64, device=device)
        self.rec_local = torch.zeros((self.rec_len,), dtype=torch.float64, device=device)
        if R > 1:
            self.max_all = torch.empty((R,), dtype=torch.float32, device=device)
            self.sum_all = torch.empty((R,), dtype=torch.float64, device=device)
            self.rec_all = torch.zeros((R, self.rec_len), dtype=torch.float64, device=device)
        else:                       # single rank: the "gathered" buffers alias the local ones
            self.max_all, self.sum_all = self.max_local, self.sum_local
            self.rec_all = self.rec_local.view(1, -1)

    # views into the local record ------------------------------------------------------------
    def state_view(self):
        """fp32 [6, n] view (row pitch n_max) of the state section of this rank's record."""
        st = self.rec_local[HEADER + self.n_max:].view(torch.float32).view(6, self.n_max)
        return st[:, :self.n]

    def stats_view(self):
        return self.rec_local[1:6]

    def local_views(self):
        """(total [1], local CDF [n], state fp32 [6, n]) inside the local record."""
        rec = self.rec_local
        return rec[0:1], rec[HEADER:HEADER + self.n], self.state_view()

    # collectives (no-ops on one rank) -------------------------------------------------------
    def gather_max(self):
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.max_all, self.max_local, group=self.pg)
        return self.max_all

    def gather_sum(self):
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.sum_all, self.sum_local, group=self.pg)
        return self.sum_all

    def gather_records(self):
        """All ranks' records, f64 [R, rec_len]: the receive buffer itself, read in place."""
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.rec_all.view(-1), self.rec_local, group=self.pg)
        return self.rec_all

    # host-side readers of a gathered buffer (tests, tools) ------------------------------------
    def unpack(self, rec_all=None):
        """(totals [R], [cdf_r [n_r]], [state_r fp32 [6, n_r]], stats [R, 5]) as tensors."""
        rec = self.rec_all if rec_all is None else rec_all
        cdfs, states = [], []
        for r, n in enumerate(self.shards):
            cdfs.append(rec[r, HEADER:HEADER + n])
            st = rec[r, HEADER + self.n_max:].view(torch.float32).view(6, self.n_max)
            states.append(st[:, :n])
        return rec[:, 0], cdfs, states, rec[:, 1:6]
