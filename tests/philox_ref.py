"""Host restatement (numpy, vectorised) of the counter-based generator used by
libsmcb: Philox4x32-10 (Salmon et al., SC'11) + the 53-bit uniform / Box-Muller
constructions of csrc/smcb_common.cuh.  TEST INFRASTRUCTURE."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
PURPOSE_NORMAL, PURPOSE_UNIFORM, PURPOSE_API = 1, 2, 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def u53(a, b):
    k = ((a >> np.uint32(5)).astype(np.uint64) << np.uint64(26)) | (b >> np.uint32(6)).astype(np.uint64)
    return k.astype(np.float64) * (1.0 / 9007199254740992.0)


def u53_open(a, b):
    k = ((a >> np.uint32(5)).astype(np.uint64) << np.uint64(26)) | (b >> np.uint32(6)).astype(np.uint64)
    return (k.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _ctr(pairs, t, w3, seed):
    pairs = np.asarray(pairs, dtype=np.uint64)
    return philox4x32_10((pairs & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                         (pairs >> np.uint64(32)).astype(np.uint32), np.uint32(t), np.uint32(w3),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def uniforms(n, t, seed, w3=PURPOSE_UNIFORM, offset=0):
    """u[i], i < n: what uniform_pair(key, (offset + i) // 2, t, w3) yields."""
    npairs = (n + 1) // 2
    r = _ctr(np.arange(npairs) + offset // 2, t, w3, seed)
    u = np.empty(2 * npairs)
    u[0::2], u[1::2] = u53(r[0], r[1]), u53(r[2], r[3])
    return u[:n]


def normals(n, t, seed, comp=0, offset=0, w3=None):
    """z[i], i < n: what normal_pair(key, (offset + i) // 2, t, comp) yields (Box-Muller)."""
    npairs = (n + 1) // 2
    if w3 is None:
        w3 = (comp << 8) | PURPOSE_NORMAL
    r = _ctr(np.arange(npairs) + offset // 2, t, w3, seed)
    u1, u2 = u53_open(r[0], r[1]), u53(r[2], r[3])
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * npairs)
    z[0::2], z[1::2] = rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
    return z[:n]
