"""The asynchronous host path ``nds_submit`` / ``nds_wait`` (one block of host memory per direction, optionally
pipelined over walker chunks on separate copy / compute streams; ``bench.py``'s ``e2e`` leg) against the plain
``nds_set_state`` -> ``nds_run`` -> ``nds_get_state`` sequence: walkers are independent and the Philox stream is
indexed by the global walker id, so every chunking must return the same bits."""
import numpy as np
import pytest

from ndsimulator_b200.engine import Engine
from tests.helpers import spec_from

pytestmark = pytest.mark.gpu

C2 = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], integrate="langevin", gamma=0.1, dt=1.0,
          temperature=300, seed=3, steps=1000)
SIDE = dict(dump=True, dump_freq=10, dump_walkers=512, dump_capacity=64, hist_n=[48, 40], hist_lo=[0.0, 0.0],
            hist_hi=[48.0, 40.0])


def start(M, nd=2):
    rng = np.random.RandomState(1)
    x0 = np.array([[22.0], [24.0]]) + rng.normal(scale=1.0, size=(nd, M))
    v0 = rng.normal(scale=0.1, size=(nd, M))
    return x0, v0


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("side_work", [False, True])
def test_submit_equals_run_for_every_chunking(prec, side_work):
    M, k = 5120 + 256 * 3 + 2, 60  # not a multiple of any chunk count; even (packed kernel)
    cfg = dict(C2, **(SIDE if side_work else {}))
    x0, v0 = start(M)
    dt = np.float32 if prec == "f32" else np.float64

    ref = Engine(spec_from(cfg, n_walkers=M, precision=prec, steps_per_launch=100).to_struct())
    ref.set_state(x0, v0)
    ref.begin()
    ref.run(k)
    ref.run(k)
    want = ref.get_state()
    want_hist = ref.get_histogram() if side_work else None

    for chunks in (1, 2, 5, 8, 16):
        eng = Engine(spec_from(cfg, n_walkers=M, precision=prec, steps_per_launch=100).to_struct())
        eng.set_state(x0, v0)
        eng.begin()
        # pass 1: state from the host block; pass 2: continue on the device state (host_in = NULL)
        xin = np.ascontiguousarray(np.concatenate([x0, v0]).astype(dt))
        out = np.zeros((2 * 2 + 5, M), dtype=dt)
        eng.submit(xin.ctypes.data, k, out.ctypes.data, chunks)
        eng.wait()
        first = out.copy()
        eng.submit(None, k, out.ctypes.data, chunks)
        eng.wait()
        assert not np.array_equal(first[:2], out[:2])
        for name, row0, n in (("x", 0, 2), ("v", 2, 2)):
            assert np.array_equal(out[row0:row0 + n].astype(np.float64), want[name]), (chunks, name)
        for j, name in enumerate(("T", "pe", "ke", "biase", "totale")):
            assert np.array_equal(out[4 + j].astype(np.float64), want[name]), (chunks, name)
        if side_work:
            assert np.array_equal(eng.get_histogram(), want_hist), chunks
            total = M * (2 * k // 10)  # one sample per walker every dump_freq steps (the few outside the box are not counted)
            assert 0.99 * total <= want_hist.sum() <= total
