"""The sampler's oracle (oracle/philox_oracle.py) against the Random123 known-answer vectors of Philox4x32-10
(Random123 kat_vectors: `philox4x32 10 ...`), and basic properties of its Box-Muller normals."""
import numpy as np

from oracle import philox_oracle as po


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = po.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in got) == want, (ctr, key, [hex(int(v)) for v in got])


def test_normals_are_shard_consistent_and_standard():
    full = po.normals(seed=99, row0=0, n=4001, d=7)
    part = po.normals(seed=99, row0=1234, n=50, d=7)
    assert np.array_equal(full[1234:1284], part)
    z = po.normals(seed=5, row0=0, n=200000, d=3).ravel()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and abs((z ** 4).mean() - 3.0) < 0.06
    assert not np.array_equal(po.normals(1, 0, 8, 4), po.normals(2, 0, 8, 4))
