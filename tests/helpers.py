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
