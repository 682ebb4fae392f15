"""CPU restatement of the counter-based Gaussian noise the fused ancestral DDPM step adds (TEST INFRASTRUCTURE ONLY).

The reference draws this noise with `torch.randn` on the device inside `DDPMScheduler.step`
(/root/reference/app.ipynb:816, no generator: an unreproducible global stream), so only the DISTRIBUTION is
reference behaviour.  The engine's stream is Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
easy as 1, 2, 3", SC'11 -- the generator behind torch's and cuRAND's device RNG) keyed by a 64-bit seed with counter
(element index lo, hi, step, 0); words 0 and 1 give two 24-bit uniforms (k + 0.5) / 2^24 and z = sqrt(-2 ln u1) cos(2 pi u2).
Pinned by the three Random123 known-answer vectors (tests/test_oracle_philox.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """uint32 arrays (broadcastable) -> four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def normal(seed: int, step: int, n: int, return_bits: bool = False):
    """float32 [n]: element i = Box-Muller of Philox(key = seed, counter = (i lo, i hi, step, 0))."""
    idx = np.arange(n, dtype=np.uint64)
    r0, r1, _, _ = philox4x32_10(idx & MASK, idx >> np.uint64(32), np.full(n, step, np.uint64), np.zeros(n, np.uint64),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = ((r0 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    u2 = ((r1 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    z = np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos(2.0 * np.pi * u2.astype(np.float64))
    z = z.astype(np.float32)
    return (z, np.stack([r0, r1], 1)) if return_bits else z
