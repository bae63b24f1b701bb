"""TEST INFRASTRUCTURE — numpy restatement of the counter-based generator behind noc_sample_rho0 (neuraloc_b200/csrc/noc_sampler.cu):
Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 philox.h) and the
Box-Muller transform.  Pinned against the Random123 known-answer vectors (tests/test_oracle_philox.py).  The reference
(src/initProb.py) draws with torch.randn on the host; there is no stream to match, only the distribution.  Only tests/ may
import this module."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4], key: uint32 [..., 2] (broadcastable) -> uint32 [..., 4]"""
    c = np.array(counter, dtype=np.uint64).copy()
    k = np.broadcast_to(np.array(key, dtype=np.uint64), c.shape[:-1] + (2,)).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[..., 0], np.uint64(M1) * c[..., 2]
        n0 = ((p1 >> np.uint64(32)) ^ c[..., 1] ^ k[..., 0]) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[..., 3] ^ k[..., 1]) & mask
        n3 = p0 & mask
        c = np.stack((n0, n1, n2, n3), axis=-1)
        k = np.stack(((k[..., 0] + np.uint64(W0)) & mask, (k[..., 1] + np.uint64(W1)) & mask), axis=-1)
    return c.astype(np.uint32)


def raw_groups(seed, group0, ngroups):
    g = np.arange(group0, group0 + ngroups, dtype=np.uint64)
    ctr = np.stack((g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), np.zeros_like(g), np.zeros_like(g)), axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    return philox4x32_10(ctr, key)


def normals(seed, row0, n, d):
    """The standard normals noc_sample_rho0 draws for rows row0 .. row0+n-1 of an [*, d] batch, in float64."""
    e0, total = row0 * d, n * d
    g0, g1 = e0 // 4, (e0 + total - 1) // 4
    w = raw_groups(seed, g0, g1 - g0 + 1).astype(np.float64)
    u = (w + 0.5) * 2.0 ** -32
    r0, r1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
    z = np.stack((r0 * np.cos(2 * np.pi * u[:, 1]), r0 * np.sin(2 * np.pi * u[:, 1]),
                  r1 * np.cos(2 * np.pi * u[:, 3]), r1 * np.sin(2 * np.pi * u[:, 3])), axis=-1).reshape(-1)
    off = e0 - 4 * g0
    return z[off:off + total].reshape(n, d)
