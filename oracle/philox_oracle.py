"""numpy restatement of the on-device synthetic volume generator (mslesions3d_b200/csrc/generate.cu) --
TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The generator follows generate_artificial_dataset.py:63-105 (uniform noise, ``randint(lo, hi) + 1`` cubes of side
``randint(smin, smax)`` at ``randint(0, dim - side)`` corners, ``+ 0.4`` and a clip per cube, binary mask) but
draws from Philox4x32-10 keyed by (seed, volume index, channel, voxel) instead of numpy's serial MT19937 stream.
This file is the independent statement of that construction the GPU parity test compares against bit for bit.
"""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10: counters uint32 arrays (broadcastable), key two Python ints -> 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & MASK32
        n1 = p1 & MASK32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & MASK32
        n3 = p0 & MASK32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def randint_ms(u, lo, hi):
    return int(lo) + int((int(u) * (int(hi) - int(lo))) >> 32)


def cubes_of(seed, idx, size, num_objects, object_size, max_cubes=64):
    """[(side, cd, ch, cw)] of volume ``idx`` (stream 1 of the generator)."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    i0 = idx & 0xFFFFFFFF
    h = philox4x32_10(0, 0, i0, 1, k0, k1)
    count = min(randint_ms(h[0], num_objects[0], num_objects[1]) + 1, max_cubes)
    out = []
    for i in range(count):
        r = philox4x32_10(1 + i, 0, i0, 1, k0, k1)
        side = randint_ms(r[0], object_size[0], object_size[1])
        out.append((side, randint_ms(r[1], 0, size[0] - side), randint_ms(r[2], 0, size[1] - side),
                    randint_ms(r[3], 0, size[2] - side)))
    return out


def volume(seed, idx, channel, size, cubes):
    """-> (raw fp32 volume (D, H, W), uint8 mask)."""
    d, h, w = size
    vox = d * h * w
    pairs = (vox + 1) // 2
    pr = np.arange(pairs, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    r = philox4x32_10(pr & MASK32, pr >> np.uint64(32), idx & 0xFFFFFFFF, (channel << 8) & 0xFFFFFFFF, k0, k1)
    u = np.empty((pairs, 2), dtype=np.float64)
    for e in range(2):
        a = (r[2 * e] >> np.uint64(5)).astype(np.float64)
        b = (r[2 * e + 1] >> np.uint64(6)).astype(np.float64)
        u[:, e] = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)
    x = u.reshape(-1)[:vox].reshape(d, h, w)
    mask = np.zeros((d, h, w), dtype=np.uint8)
    for side, cd, ch, cw in cubes:
        sl = (slice(cd, cd + side), slice(ch, ch + side), slice(cw, cw + side))
        x[sl] = np.clip(x[sl] + 0.4, 0.0, 1.0)
        mask[sl] = 1
    return x.astype(np.float32), mask
