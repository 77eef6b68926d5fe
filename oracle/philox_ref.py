"""numpy restatement of the counter-based noise the CUDA kernels generate (TEST
INFRASTRUCTURE, not product).

The reference draws its Laplace and Gumbel noise from torch's global generators
(models.py:74 `self.noiser.sample`, models.py:77 `F.gumbel_softmax` -> `exponential_()`),
which are sequential and device-specific.  The product path replaces them by Philox4x32-10
(Salmon et al., SC'11; the same round function cuRAND and torch-CUDA use) keyed so that
every (sample row, column, pass) has its own counter: noise is independent of batch
partitioning and of the GPU count.  This file defines that mapping bit-for-bit; the CUDA
side (`csrc/philox.cuh`) must produce the same uint32 words, checked in
`tests/test_gpu_kernels.py::test_perturb_gate_philox_matches_oracle_noise`.

Counter / key layout (one Philox call yields the four words of four consecutive columns):
    counter = (col // 4, row_global, stream, offset)      key = (seed_lo, seed_hi)
    stream 0: Laplace bits, stream 1: Gumbel plane 0 (logit w), stream 2: Gumbel plane 1 (1-w)
    word j of the call belongs to column 4*(col//4) + j.

Float transforms (each exact in fp32 before the log):
    Laplace(0,1):  sign = r>>31, v = ((r & 0x7FFFFF) + 0.5) * 2^-23,  x = -+log(v)
    Gumbel(0,1):   v = ((r>>9) + 0.5) * 2^-23,  E = -log(v) ~ Exp(1),  g = -log(E)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_LAPLACE = 0
STREAM_GUMBEL0 = 1
STREAM_GUMBEL1 = 2


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10.  c* are uint32 arrays (broadcastable), k* python ints."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def words(seed: int, offset: int, stream: int, row0: int, B: int, D: int) -> np.ndarray:
    """uint32 [B,D] random words for rows row0..row0+B-1, all D columns (D % 4 == 0)."""
    assert D % 4 == 0
    cols4 = np.arange(D // 4, dtype=np.uint32)[None, :]
    rows = (np.arange(B, dtype=np.uint64) + np.uint64(row0)).astype(np.uint32)[:, None]
    out = philox4x32_10(cols4, rows, np.uint32(stream), np.uint32(offset & 0xFFFFFFFF),
                        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(out, axis=-1).reshape(B, D)


def laplace(seed: int, offset: int, row0: int, B: int, D: int) -> np.ndarray:
    r = words(seed, offset, STREAM_LAPLACE, row0, B, D)
    v = ((r & np.uint32(0x7FFFFF)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    mag = -np.log(v.astype(np.float64)).astype(np.float32)
    return np.where((r >> np.uint32(31)) != 0, -mag, mag).astype(np.float32)


def gumbel(seed: int, offset: int, row0: int, B: int, D: int) -> np.ndarray:
    """[2,B,D] Gumbel(0,1) noise for the two gate planes."""
    out = []
    for stream in (STREAM_GUMBEL0, STREAM_GUMBEL1):
        r = words(seed, offset, stream, row0, B, D)
        v = ((r >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
        e = -np.log(v.astype(np.float64))
        out.append((-np.log(e)).astype(np.float32))
    return np.stack(out, axis=0)
