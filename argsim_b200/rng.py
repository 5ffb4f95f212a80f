"""numpy restatement of the device/host Philox4x32-10 streams (csrc/philox.h) so that keep-masks
and eps drawn by the library can be reproduced for parity runs.  Counter = (global row, position,
stream, seed_hi); key = (seed_lo, step)."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
STREAM_KEEP, STREAM_EPS = 0, 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, np.uint32) for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def keep_mask(b, T, rate_keepwd, seed, step, row0=0, rows=None):
    """(b,T) uint8 mask the library draws for word dropout (model.py:94) when none is injected; rows = global row
    index of every row (argsim_set_global_rows), default row0 + i."""
    gid = np.arange(b, dtype=np.int64) + row0 if rows is None else np.asarray(rows, np.int64)
    rows = gid[:, None].astype(np.uint32) + np.zeros((1, T), np.uint32)
    pos = np.zeros((b, 1), np.uint32) + np.arange(T, dtype=np.uint32)[None, :]
    c0, _, _, _ = philox4x32_10(rows, pos, np.uint32(STREAM_KEEP), np.uint32((seed >> 32) & 0xFFFFFFFF),
                                seed & 0xFFFFFFFF, step & 0xFFFFFFFF)
    u = (c0 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (u < np.float32(rate_keepwd)).astype(np.uint8)
