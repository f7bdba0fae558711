"""GPU: the Tracker's sharded PF path, emulated on ONE device by running R Tracker-like shards
sequentially through the same kernels, must equal the single-shard run bit for bit."""
import numpy as np
import pytest
import torch

from gpu_util import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R", [2, 4])
def test_sharded_pf_step_is_rank_invariant(R):
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    n, D, tau, seed, t = 1024, 192, 0.1, 3, 5
    g = torch.Generator(device="cpu").manual_seed(1)
    tmpl = torch.randn(D, generator=g).to(dev)
    u = ops.resample_uniform(seed, t)

    def run(R):
        nl = n // R
        state = torch.zeros((6, n), device=dev)
        emb = torch.empty((n, D), device=dev)
        for r in range(R):                       # per-rank init + propagate with global ids
            st = torch.zeros((6, nl), device=dev)
            ops.init_particles(st, 100.0, 80.0, 4.0, seed, 0, r * nl)
            ops.propagate(st, 2.0, 1.0, 0.01, seed, t, 0, r * nl)
            state[:, r * nl:(r + 1) * nl] = st
        # embeddings are a deterministic function of the particle (stand-in for crop + ViT)
        emb.copy_(tmpl[None] + 0.5 * torch.sin(state[0][:, None] * torch.arange(1, D + 1, device=dev)))
        logw = torch.empty(n, device=dev); mx = torch.empty(R, device=dev)
        sm = torch.empty(R, dtype=torch.float64, device=dev)
        w = torch.empty(n, device=dev); cdf = torch.empty(n, dtype=torch.float64, device=dev)
        tot = torch.empty(R, dtype=torch.float64, device=dev)
        stats = torch.zeros(5, dtype=torch.float64, device=dev)
        for r in range(R):
            sl = slice(r * nl, (r + 1) * nl)
            ops.loglik(emb[sl], tmpl, tau, logw[sl], mx[r:r + 1])
        for r in range(R):
            ops.sumexp(logw[r * nl:(r + 1) * nl], mx, sm[r:r + 1])
        for r in range(R):
            sl = slice(r * nl, (r + 1) * nl)
            ops.weights(logw[sl], mx, sm, state[:, sl].contiguous(), w[sl], cdf[sl], tot[r:r + 1], stats)
        state_all = state.view(6, R, nl).permute(1, 0, 2).contiguous()
        out = torch.empty((6, n), device=dev); idx = torch.empty(n, dtype=torch.int32, device=dev)
        for r in range(R):
            o = torch.empty((6, nl), device=dev)
            ops.resample(cdf, tot, nl, r * nl, u, state_all, o, idx[r * nl:(r + 1) * nl])
            out[:, r * nl:(r + 1) * nl] = o
        return state.clone(), w, idx, out

    s1, w1, i1, o1 = run(1)
    sR, wR, iR, oR = run(R)
    assert torch.equal(s1, sR) and torch.equal(w1, wR) and torch.equal(i1, iR) and torch.equal(o1, oR)


def test_tracker_on_a_non_current_device():
    """a tracker built on cuda:1 while cuda:0 is current must launch on cuda:1 (libvpf
    takes torch's current stream and keys its per-device state on cudaGetDevice)."""
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    frames, gt = make_clip("box", 224, 224, 4, seed=0, target=(48, 48))
    a = Tracker("vit_tiny_patch16_224", 64, seed=3, device="cuda:0")
    b = Tracker("vit_tiny_patch16_224", 64, seed=3, device="cuda:1")     # cuda:0 stays current
    a.init(frames[0], gt[0])
    b.init(frames[0], gt[0])
    assert b.particles.device.index == 1 and torch.cuda.current_device() == 0
    for t in (1, 2, 3):
        sa, sb = a.step(frames[t]), b.step(frames[t])
        assert sa == sb and torch.equal(a.particles.cpu(), b.particles.cpu())
    assert torch.cuda.current_device() == 0


def test_graph_mode_survives_an_interleaved_hook_step():
    """a hooked (eager) step between graph replays must not leave the captured graph
    with stale pointers."""
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    frames, gt = make_clip("box", 224, 224, 8, seed=1, target=(48, 48))
    eager = Tracker("vit_tiny_patch16_224", 128, seed=5)
    graph = Tracker("vit_tiny_patch16_224", 128, seed=5, cuda_graph=True)
    eager.init(frames[0], gt[0])
    graph.init(frames[0], gt[0])
    for t in range(1, 8):
        hook = {"_u": 0.25} if t in (3, 6) else {}
        a, b = eager.step(frames[t], **hook), graph.step(frames[t], **hook)
        assert torch.equal(eager.particles, graph.particles), t
        assert torch.equal(eager.idx, graph.idx) and a == b
    assert graph._graph is not None


def test_particle_count_bound_is_enforced():
    from vitparticlefiltertracker_b200 import Tracker
    cuda()
    with pytest.raises(ValueError, match="n_particles"):
        Tracker("vit_tiny_patch16_224", (1 << 17) + 1)
