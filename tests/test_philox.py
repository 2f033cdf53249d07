"""Known-answer vectors of Philox4x32-10 (Random123 kat_vectors) for the NumPy mirror of the
device RNG; the GPU tests then compare the device draws with this mirror."""
import numpy as np

from oracle import philox_ref as ph


def _run(ctr, key):
    out = ph.philox4x32_10(*[np.uint64(c) for c in ctr], key[0], key[1])
    return tuple(int(o) for o in out)


def test_random123_kat():
    assert _run((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert _run((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert _run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_distributions():
    z = ph.half_normal(seed=7, env=3, episode=1, c1=5, elems=np.arange(200000))
    assert z.min() >= 0
    assert abs(z.mean() - np.sqrt(2 / np.pi)) < 5e-3           # E|N(0,1)|
    assert abs((z ** 2).mean() - 1.0) < 1e-2
    hours = [ph.start_time(0, e, 1, 0, 9, 20) for e in range(4000)]
    h = np.array(hours)
    assert h[:, 0].min() == 0 and h[:, 0].max() == 23
    assert h[:, 1].min() == 0 and h[:, 1].max() == 8
    assert h[:, 2].min() == 0 and h[:, 2].max() == 19
    a = ph.uniform_action(0, 1, 1, 0, 10000, -0.8, 0.8)
    assert a.min() >= -0.8 and a.max() < 0.8 and abs(a.mean()) < 0.02
