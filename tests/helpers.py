"""Shared test helpers: golden fixtures, config plumbing, comparison metrics."""
import json
import os

import numpy as np

from ndsimulator_b200 import config as nds_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def spec_from(cfg, **extra):
    c = dict(cfg)
    c.update(extra)
    c.setdefault("dump", False)
    return nds_config.parse(c)


def relerr(a, b, floor=1e-12):
    """max |a-b| / max(|b|, floor * scale) -- relative error with an absolute floor at the data scale."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * scale + 1e-300))) if b.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| : error relative to the magnitude of the whole vector."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


# ---------------------------------------------------------------------------- Philox4x32-10 (numpy)
def philox4x32_10(ctr, key):
    """Reference implementation of Philox4x32-10 (Salmon et al., SC'11; Random123) on Python ints."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(v) & 0xFFFFFFFF for v in ctr]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c


# Random123 known-answer vectors (kat_vectors: philox4x32 10)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox_normal_fields(seed, walker, call, domain=0):
    """The three (radius field, angle field) pairs of 21 bits of one Philox call and the words of the extension call
    (same counter, top bit of the call word set) the fp64 engine adds (ndsimulator_b200/csrc/nds_philox.cuh)."""
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) & 0xFFFFFFFF) ^ domain
    ctr = (call & 0xFFFFFFFF, call >> 32, walker & 0xFFFFFFFF, walker >> 32)
    r = philox4x32_10(ctr, (k0, k1))
    e = philox4x32_10((ctr[0], ctr[1] | 0x80000000, ctr[2], ctr[3]), (k0, k1))
    fields = ((r[0] >> 11, r[1] >> 11), (r[2] >> 11, r[3] >> 11),
              ((r[0] & 0x7FF) | ((r[1] & 0x3FF) << 11), (r[2] & 0x7FF) | ((r[3] & 0x3FF) << 11)))
    ext = ((e[0], e[3] >> 21, 11), (e[1], (e[3] >> 10) & 0x7FF, 11), (e[2], e[3] & 0x3FF, 10))
    return fields, ext


def philox_normals_f64(seed, walker, n_normals, domain=0):
    """The engine's fp64 normal stream of one walker: six normals per call, three Box-Muller pairs of a 53-bit radius
    variate u1 = (a 2^32 + e + 1/2) 2^-53 and a 32 / 32 / 31-bit angle (nds_philox.cuh normals6<double>)."""
    out = []
    call = 0
    while len(out) < n_normals:
        fields, ext = philox_normal_fields(seed, walker, call, domain)
        for (a, b), (ea, eb, bits) in zip(fields, ext):
            u1 = (float((a << 32) | ea) + 0.5) * 2.0 ** -53
            u2 = float((b << bits) | eb) * 2.0 ** -(21 + bits)
            rad = np.sqrt(-2.0 * np.log(u1))
            out += [rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)]
        call += 1
    return np.array(out[:n_normals])


def philox_normals_f32(seed, walker, n_normals, domain=0):
    """The fp32 engines' stream in exact arithmetic: 21-bit fields, u1 = (a + 1/2) 2^-21 (cell midpoint), angle
    2 pi b 2^-21.  The device evaluates this with MUFU approximations (lg2 / sqrt / sin / cos .approx)."""
    out = []
    call = 0
    while len(out) < n_normals:
        fields, _ = philox_normal_fields(seed, walker, call, domain)
        for a, b in fields:
            rad = np.sqrt(-2.0 * np.log((a + 0.5) * 2.0 ** -21))
            out += [rad * np.cos(2 * np.pi * b * 2.0 ** -21), rad * np.sin(2 * np.pi * b * 2.0 ** -21)]
        call += 1
    return np.array(out[:n_normals])


def fp32_tail_probability(t):
    """P(|z| > t) of ONE normal of the fp32 scheme in exact arithmetic: the radius takes the 2^21 values r_n =
    sqrt(-2 ln((n + 1/2) 2^-21)), the angle is (to 2^-21) uniform, z = r cos / sin(angle)."""
    n_max = int(np.floor(2.0 ** 21 * np.exp(-t * t / 2.0) + 1))
    n = np.arange(0, min(n_max, 2 ** 21))
    r = np.sqrt(-2.0 * np.log((n + 0.5) * 2.0 ** -21))
    r = r[r > t]
    return float(np.sum(2.0 / np.pi * np.arccos(t / r)) * 2.0 ** -21)
