"""oracle/philox.py — TEST INFRASTRUCTURE ONLY.

numpy restatement of the counter-based eps generator of csrc/bayes.cu (bem_bayes_sample with eps == NULL):
Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), key = (seed lo, seed hi),
counter = (block lo, block hi, sample, stream_id), block = element_index // 4; the 4 outputs feed two Box-Muller pairs and
element i takes output i % 4. The reference itself draws eps with torch's `normal_()` (basicsr/bayesian/conv.py:107);
this generator exists so that Monte-Carlo samples are reproducible independent of how they are sharded over GPUs.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, np.uint32) for v in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def normal(numel, seed, stream_id, sample):
    """eps[0:numel] for one (seed, stream_id, sample), float32."""
    nblk = (numel + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    c0 = (blk & MASK).astype(np.uint32)
    c1 = (blk >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(nblk, sample & 0xFFFFFFFF, np.uint32)
    c3 = np.full(nblk, stream_id & 0xFFFFFFFF, np.uint32)
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    k = np.float32(2.0 ** -24)
    u = [((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * k for r in (r0, r1, r2, r3)]
    ra = np.sqrt(np.float32(-2.0) * np.log(u[0])).astype(np.float32)
    rb = np.sqrt(np.float32(-2.0) * np.log(u[2])).astype(np.float32)
    ta = np.float32(6.283185307179586) * u[1]
    tb = np.float32(6.283185307179586) * u[3]
    out = np.stack([ra * np.cos(ta), ra * np.sin(ta), rb * np.cos(tb), rb * np.sin(tb)], axis=1).astype(np.float32)
    return out.reshape(-1)[:numel]
