"""Known-answer pins of oracle/philox.py: the three philox4x32-10 vectors of Random123's kat_vectors file (the
generator's published test vectors), plus distribution checks of the Gaussian transform."""
import numpy as np

from oracle import philox as P

KAT = [  # counter, key, expected (Random123 kat_vectors: "philox4x32 10 ...")
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, exp in KAT:
        got = tuple(int(v) for v in P.philox4x32_10(*ctr, *key))
        assert got == exp, (ctr, key, [hex(v) for v in got])


def test_normal_is_a_pure_function_and_gaussian():
    z = P.normal(1234567890123, 7, 1 << 18)
    assert np.array_equal(z, P.normal(1234567890123, 7, 1 << 18))
    assert np.array_equal(z[:1000], P.normal(1234567890123, 7, 1000))          # element i does not depend on n
    assert not np.array_equal(z[:1000], P.normal(1234567890123, 8, 1000))      # step and seed select the stream
    assert not np.array_equal(z[:1000], P.normal(1234567890124, 7, 1000))
    assert np.isfinite(z).all()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3.0) < 0.08    # skewness, kurtosis
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.01                          # neighbouring elements uncorrelated
