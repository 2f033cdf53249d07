"""ORACLE - TEST INFRASTRUCTURE ONLY. NumPy mirror of the counter-based RNG the CUDA env uses
(``mapdn_b200/csrc/philox.cuh``): Philox4x32-10 (Salmon et al., SC'11) keyed by
(seed, env, episode, step, element).

The reference draws from the process-global ``np.random`` stream
(reference voltage_control_env.py:49,384,389,398,498,503,508,337), which couples episode
sampling, noise and the replay buffer; trajectory-level reproduction of it is impossible by
construction. The *distributions* are preserved: |N(0,1)| noise, U{0..n-1} start
indices, U[low, high) reset actions."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_NOISE, STREAM_TIME, STREAM_ACTION = 0, 1, 2
RESET_FLAG = 0x80000000


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over equal-shape uint64 arrays holding 32-bit values. Returns 4 uint64 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(hi, lo):
    """(0,1) double from two 32-bit words: ((hi<<32|lo)>>11 + 0.5) * 2^-53."""
    bits = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (bits.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _c3(episode, stream):
    return (int(episode) * 8 + stream) & 0xFFFFFFFF


def half_normal(seed, env, episode, c1, elems):
    """|N(0,1)| for element indices ``elems``: elements (2m, 2m+1) share one Philox call (Box-Muller pair:
    cosine branch for the even element, sine branch for the odd one)."""
    elems = np.asarray(elems, np.uint64)
    x0, x1, x2, x3 = philox4x32_10(elems >> np.uint64(1), np.uint64(c1), np.uint64(env),
                                   np.uint64(_c3(episode, STREAM_NOISE)), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1, u2 = _u53(x0, x1), _u53(x2, x3)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return np.abs(rad * np.where((elems & np.uint64(1)) == 1, np.sin(ang), np.cos(ang)))


def start_time(seed, env, episode, attempt, n_day_choices, steps_per_hour):
    """(hour, day, interval) ~ U{0..23}, U{0..n_day_choices-1}, U{0..steps_per_hour-1}
    via the unbiased-to-2^-32 multiply-high map."""
    x0, x1, x2, _ = philox4x32_10(np.uint64(0), np.uint64(attempt), np.uint64(env),
                                  np.uint64(_c3(episode, STREAM_TIME)),
                                  seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    hour = int((int(x0) * 24) >> 32)
    day = int((int(x1) * int(n_day_choices)) >> 32)
    interval = int((int(x2) * int(steps_per_hour)) >> 32)
    return hour, day, interval


def uniform_action(seed, env, episode, attempt, n, low, high):
    j = np.arange(n, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(j, np.uint64(attempt), np.uint64(env), np.uint64(_c3(episode, STREAM_ACTION)),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return low + (high - low) * _u53(x0, x1)
