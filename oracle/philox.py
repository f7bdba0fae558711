"""Philox4x32-10 counter RNG (numpy, vectorised) -- oracle, test infrastructure only.

Follows SURVEY.md Appendix A, D7 (the reference, /root/reference/README.md:8, says only
"probabilistic algorithms"; no RNG is named, so Random123's Philox4x32-10 is restated from its
published constants and pinned by the Random123 known-answer vectors in tests/).

Counter layout used throughout the tracker:
    ctr = (global_particle_id, frame_idx, stream_id, draw)      key = (seed_lo, seed_hi)
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 array-likes of uint32 (broadcastable); key: 2 uint32 scalars/arrays.

    Returns a tuple of 4 uint32 arrays.  Ten rounds, key bumped by the Weyl constants between
    rounds (Random123 `philox4x32_R(10, ...)`).
    """
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in ctr]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = np.asarray(key[0], dtype=np.uint32)
    k1 = np.asarray(key[1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = c0.astype(np.uint64) * PHILOX_M0
            p1 = c2.astype(np.uint64) * PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return c0, c1, c2, c3


def _split_seed(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)


def box_muller(ra, rb):
    """Two uint32 arrays -> two fp32 standard normals (D7, revised by D14).

    u1 = ((ra >> 8) + 1) * 2^-24  in (0, 1]   (exactly representable in fp32)
    u2 =  (rb >> 8)      * 2^-24  in [0, 1)
    z0 = fp32(sqrt(-2 ln u1) * cos(2 pi u2)),  z1 = fp32(sqrt(-2 ln u1) * sin(2 pi u2)).

    D14: the transcendentals are evaluated in fp64 and the product is rounded ONCE to fp32.  An
    fp64 libm result (error < 1-2 ulp of fp64) rounds to the same fp32 value whatever library
    produced it (up to a ~2^-28 chance per value of straddling a rounding boundary), so the CPU
    oracle and the GPU agree bit for bit instead of "to fp32 transcendental ulp".
    """
    u1 = ((ra >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (rb >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 6.283185307179586 * u2
    return (rad * np.cos(ang)).astype(np.float32), (rad * np.sin(ang)).astype(np.float32)


def philox_normals(ids, frame_idx: int, stream_id: int, seed: int):
    """Six fp32 normals per particle id: draws 0 and 1 of the (id, frame, stream) counter.

    Returns float32 array [6, len(ids)]; rows 0-3 come from draw 0 (pairs (r0,r1), (r2,r3)),
    rows 4-5 from draw 1 (pair (r0,r1)).
    """
    ids = np.asarray(ids, dtype=np.uint32)
    key = _split_seed(seed)
    f = np.uint32(frame_idx)
    s = np.uint32(stream_id)
    a = philox4x32_10((ids, f, s, np.uint32(0)), key)
    b = philox4x32_10((ids, f, s, np.uint32(1)), key)
    z0, z1 = box_muller(a[0], a[1])
    z2, z3 = box_muller(a[2], a[3])
    z4, z5 = box_muller(b[0], b[1])
    return np.stack([z0, z1, z2, z3, z4, z5]).astype(np.float32)


def philox_uniform53(frame_idx: int, stream_id: int, seed: int) -> float:
    """The one fp64 uniform per frame used by systematic resampling (D6).

    counter (0xFFFFFFFF, frame_idx, stream_id, 2);  u = ((r0 << 21) ^ (r1 >> 11)) * 2^-53.
    """
    key = _split_seed(seed)
    r = philox4x32_10((np.uint32(0xFFFFFFFF), np.uint32(frame_idx), np.uint32(stream_id),
                       np.uint32(2)), key)
    hi = int(r[0])
    lo = int(r[1])
    v = ((hi << 21) ^ (lo >> 11)) & ((1 << 53) - 1)
    return float(v) * (2.0 ** -53)
