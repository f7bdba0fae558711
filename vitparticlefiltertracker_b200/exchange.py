"""Host-side exchange protocol of the particle-sharded multi-GPU path.

Device-agnostic torch code: NCCL on GPUs in the product, gloo on CPU in tests
(tests/test_exchange_gloo.py, world_size 2).  Three tiny all-gathers per frame; every exchanged
quantity is exact (a max, or an fp64 sum of grid-snapped terms), so the result does not depend
on the number of ranks.

Per-rank record, fp64 slots:  [ total (1) | local inclusive CDF (n) | state fp32 [6][n] (3n) ]
"""
from __future__ import annotations

import torch


def shard_range(n_global: int, rank: int, world: int):
    """Rank r owns global particle ids [r*n, (r+1)*n), n = n_global / world."""
    if n_global % world:
        raise ValueError("n_particles must be divisible by the number of ranks")
    n = n_global // world
    return rank * n, n


class Exchange:
    def __init__(self, process_group, n_local: int, device):
        import torch.distributed as dist
        self.dist = dist
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.rank = dist.get_rank(process_group) if process_group is not None else 0
        self.n = n_local
        R, n = self.world, n_local
        self.rec_len = 1 + n + 3 * n
        self.max_local = torch.empty((1,), dtype=torch.float32, device=device)
        self.sum_local = torch.empty((1,), dtype=torch.float64, device=device)
        self.rec_local = torch.empty((self.rec_len,), dtype=torch.float64, device=device)
        if R > 1:
            self.max_all = torch.empty((R,), dtype=torch.float32, device=device)
            self.sum_all = torch.empty((R,), dtype=torch.float64, device=device)
            self.rec_all = torch.empty((R, self.rec_len), dtype=torch.float64, device=device)
        else:                       # single rank: the "gathered" buffers alias the local ones
            self.max_all, self.sum_all = self.max_local, self.sum_local
            self.rec_all = self.rec_local.view(1, -1)

    # views into the local record ------------------------------------------------------------
    def local_views(self):
        n = self.n
        rec = self.rec_local
        return rec[0:1], rec[1:1 + n], rec[1 + n:].view(torch.float32).view(6, n)

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
        """Returns (totals [R], cdf_all [R,n], state_all [R,6,n]) as contiguous tensors."""
        n, R = self.n, self.world
        if R > 1:
            self.dist.all_gather_into_tensor(self.rec_all.view(-1), self.rec_local, group=self.pg)
            totals = self.rec_all[:, 0].contiguous()
            cdf_all = self.rec_all[:, 1:1 + n].contiguous()
            state_all = self.rec_all[:, 1 + n:].contiguous().view(torch.float32).view(R, 6, n)
            return totals, cdf_all, state_all
        total, cdf, state = self.local_views()
        return total, cdf.view(1, n), state.view(1, 6, n)

    def reduce_stats(self, stats):
        if self.world > 1:
            self.dist.all_reduce(stats, group=self.pg)
        return stats
