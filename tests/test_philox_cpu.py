"""oracle/philox.py (the numpy statement of the counter RNG that oo_rng_fill / oo_rng_fill_rows / K2 implement) against the
published known-answer vectors of Philox4x32-10 (Random123, kat_vectors), and the layout of the two streams built on it."""
import numpy as np

import philox


def block(c, k):
    r = philox.philox4x32_10(*[np.array([x], dtype=np.uint32) for x in c], k[0], k[1])
    return [int(x[0]) for x in r]


def test_philox4x32_10_known_answers():
    assert block([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert block([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert block([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_element_stream_layout():
    """oo_rng_fill: element i = word i % 4 of philox(counter = (i // 4, 0, object id, frame), key = seed)."""
    seed, frame, oid = 0x0123456789ABCDEF, 17, 42
    w = philox.words(seed, frame, oid, 10)
    for i in (0, 3, 4, 9):
        assert int(w[i]) == block([i // 4, 0, oid, frame], [seed & 0xFFFFFFFF, seed >> 32])[i % 4]
    u = philox.uniform(seed, frame, oid, 4097)
    assert u.dtype == np.float32 and 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02


def test_ray_blocked_stream_layout():
    """oo_rng_fill_rows / K2 counter mode: word w of row r = word w % 4 of philox(counter = (r, w // 4, object id, frame))."""
    seed, frame, oid = 77, 8 * 5 + 1, 9
    W = philox.words_rows(seed, frame, oid, 50, 14)
    assert W.shape == (50, 16)
    for r, w in ((0, 0), (7, 5), (49, 13), (3, 15)):
        assert int(W[r, w]) == block([r, w // 4, oid, frame], [seed & 0xFFFFFFFF, seed >> 32])[w % 4]
    U = philox.uniform_rows(seed, frame, oid, 2000, 14)
    assert U.shape == (2000, 14) and abs(U.mean() - 0.5) < 0.01
    assert abs(np.corrcoef(U[:, 0], U[:, 1])[0, 1]) < 0.08 and abs(np.corrcoef(U[:-1, 4], U[1:, 4])[0, 1]) < 0.08
    N = philox.normal_rows(seed, frame, oid, 4000, 14, 0.1 / 3)
    assert abs(N.std() - 0.1 / 3) < 1e-3 and abs(N.mean()) < 2e-3
    # the pairs of a row are the cos / sin branches of one Box-Muller draw: n0^2 + n1^2 = -2 std^2 log(u0)
    w = philox.words_rows(seed, frame, oid, 4000, 14)
    u0 = ((w[:, 0] >> np.uint32(8)).astype(np.float64) + 1) * 2.0 ** -24
    np.testing.assert_allclose(N[:, 0].astype(np.float64) ** 2 + N[:, 1].astype(np.float64) ** 2,
                               -2 * (0.1 / 3) ** 2 * np.log(u0), rtol=1e-4, atol=1e-9)
