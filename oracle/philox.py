"""Philox4x32-10 counter RNG in numpy (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) --
the published algorithm oo_rng_fill implements.  TEST INFRASTRUCTURE ONLY.
Pinned: tests/test_philox_cpu.py checks it against the Random123 known-answer vectors of Philox4x32-10; the GPU tests
check the kernels against it word for word.
Element i of object `oid` in frame `frame` = word (i % 4) of philox(counter=(i//4 lo, i//4 hi, oid, frame), key=seed)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def words(seed, frame, oid, n):
    q = np.arange((n + 3) // 4, dtype=np.uint64)
    r = philox4x32_10((q & np.uint64(0xFFFFFFFF)).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                      np.full(q.shape, oid, np.uint32), np.full(q.shape, frame, np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(r, axis=1).reshape(-1)[:4 * ((n + 3) // 4)]


def uniform(seed, frame, oid, n):
    w = words(seed, frame, oid, n)
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))[:n]


def normal(seed, frame, oid, n, std):
    w = words(seed, frame, oid, n).reshape(-1, 4)
    f = lambda x: (x >> np.uint32(8)).astype(np.float32)
    u0, u1 = (f(w[:, 0]) + 1) * np.float32(2.0 ** -24), f(w[:, 1]) * np.float32(2.0 ** -24)
    u2, u3 = (f(w[:, 2]) + 1) * np.float32(2.0 ** -24), f(w[:, 3]) * np.float32(2.0 ** -24)
    ra = np.sqrt(-2 * np.log(u0.astype(np.float64))) * std
    rb = np.sqrt(-2 * np.log(u2.astype(np.float64))) * std
    out = np.stack([ra * np.cos(2 * np.pi * u1.astype(np.float64)), ra * np.sin(2 * np.pi * u1.astype(np.float64)),
                    rb * np.cos(2 * np.pi * u3.astype(np.float64)), rb * np.sin(2 * np.pi * u3.astype(np.float64))], axis=1)
    return out.reshape(-1)[:n].astype(np.float32)


def words_rows(seed, frame, oid, n_rows, row_words):
    """Ray-blocked stream (oo_rng_fill_rows): word w of row r = word (w % 4) of philox(counter=(r, w // 4, oid, frame))."""
    nblk = (row_words + 3) // 4
    rows = np.repeat(np.arange(n_rows, dtype=np.uint32), nblk)
    blk = np.tile(np.arange(nblk, dtype=np.uint32), n_rows)
    r = philox4x32_10(rows, blk, np.full(rows.shape, oid, np.uint32), np.full(rows.shape, frame, np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(r, axis=1).reshape(n_rows, 4 * nblk)


def uniform_rows(seed, frame, oid, n_rows, row_words):
    w = words_rows(seed, frame, oid, n_rows, row_words)
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))[:, :row_words]


def normal_rows(seed, frame, oid, n_rows, row_words, std):
    w = words_rows(seed, frame, oid, n_rows, row_words).reshape(n_rows, -1, 2)
    f = lambda x: (x >> np.uint32(8)).astype(np.float32)
    u0, u1 = (f(w[..., 0]) + 1) * np.float32(2.0 ** -24), f(w[..., 1]) * np.float32(2.0 ** -24)
    rad = np.sqrt(-2 * np.log(u0.astype(np.float64))) * std
    out = np.stack([rad * np.cos(2 * np.pi * u1.astype(np.float64)), rad * np.sin(2 * np.pi * u1.astype(np.float64))], axis=-1)
    return out.reshape(n_rows, -1)[:, :row_words].astype(np.float32)
