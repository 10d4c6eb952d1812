"""Counter-based sampler shared (bit-exactly) by the oracle and the CUDA path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws float16 normals / uniforms from JAX's threefry generator
(``solvers/ScaSML.py:180-181,190,228-229``; ``solvers/ScaSML_full_history.py:
103-104,110,144,149``).  Threefry streams are JAX-version dependent and not
reproducible here, but their *aliasing structure* is part of the reference's
semantics: a value depends only on (key, flat element
index), the terminal draws re-use one fixed key in every call, and the
full-history variants use that same fixed key for every draw.

Restatement: Philox4x32-10, ``value = F(key, flat index)``:
  * counter = flat_index // 8 (64-bit, in the two low counter words),
    16-bit chunk ``flat_index % 8`` of the 128 output bits (word j//2, low half
    first);
  * normal  = inverse-CDF of p = (chunk + 0.5) / 65536, evaluated once on the
    host in float64 (scipy ``ndtri``) and rounded to float16 -> a 32768-entry
    half-table plus a sign (antisymmetric), so oracle and device agree by
    construction;
  * uniform = (chunk >> 5) / 2048 in [0, 1): an 11-bit value, exactly
    representable in float16, monotone in the same chunk as the normal
    (so ``tau ~ Phi(normal)`` like the reference's shared-key draws).
Keys: (k0, k1) with k0 = stream id (0 for the fixed key, the running split
counter for quadrature step draws) and k1 = (domain << 31) | (seed & 0x7fffffff),
domain 0 = fixed key, 1 = step keys.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

DOMAIN_FIXED = 0
DOMAIN_STEP = 1


def make_key(stream, domain, seed=0):
    k0 = int(stream) & 0xFFFFFFFF
    k1 = ((int(domain) & 1) << 31) | (int(seed) & 0x7FFFFFFF)
    return k0, k1


def philox4x32_10(ctr_lo, ctr_hi, k0, k1):
    """Vectorised Philox4x32-10. ctr_lo/ctr_hi: uint64 arrays holding 32-bit words."""
    c0 = ctr_lo.astype(np.uint64) & MASK32
    c1 = ctr_hi.astype(np.uint64) & MASK32
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0 = int(k0)
    k1 = int(k1)
    for r in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def chunks16(key, start, count):
    """16-bit chunks for flat indices start .. start+count-1 (uint32 array)."""
    k0, k1 = key
    start = int(start)
    count = int(count)
    if count == 0:
        return np.zeros(0, dtype=np.uint32)
    first_blk = start // 8
    last_blk = (start + count - 1) // 8
    blk = np.arange(first_blk, last_blk + 1, dtype=np.uint64)
    w = philox4x32_10(blk & MASK32, blk >> np.uint64(32), k0, k1)
    words = np.stack(w, axis=1)                              # [nblk, 4]
    lo = (words & np.uint64(0xFFFF)).astype(np.uint32)
    hi = ((words >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint32)
    ch = np.stack([lo, hi], axis=2).reshape(-1)              # [nblk*8]
    off = start - first_blk * 8
    return ch[off:off + count]


_HALF_TABLE = None


def normal_half_table():
    """T[i] = float16(ndtri(0.5 + (i + 0.5) / 65536)), i in [0, 32768)."""
    global _HALF_TABLE
    if _HALF_TABLE is None:
        from scipy.special import ndtri
        i = np.arange(32768, dtype=np.float64)
        _HALF_TABLE = ndtri(0.5 + (i + 0.5) / 65536.0).astype(np.float16)
    return _HALF_TABLE


def chunk_to_normal(ch):
    t = normal_half_table().astype(np.float64)
    ch = ch.astype(np.int64)
    neg = ch < 32768
    idx = np.where(neg, 32767 - ch, ch - 32768)
    v = t[idx]
    return np.where(neg, -v, v)


def chunk_to_uniform(ch):
    return (ch >> 5).astype(np.float64) / 2048.0


def normals(key, start, count):
    """float16-valued standard normals (as float64) for flat indices start.."""
    return chunk_to_normal(chunks16(key, start, count))


def uniforms(key, start, count):
    return chunk_to_uniform(chunks16(key, start, count))
