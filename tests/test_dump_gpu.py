"""Device-side dump rows (io/dump.py:76-102 content) and the in-kernel colvar histogram: float rows for fp32 engines,
frames from the packed (two walkers per thread) kernel, the split launch for subset dumps ("CV
histogram on-device + full state of walkers 0..4095 every 10 steps")."""
import numpy as np
import pytest

from ndsimulator_b200.engine import Engine
from tests.helpers import spec_from

pytestmark = pytest.mark.gpu

C2 = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md", integrate="langevin", gamma=0.1,
          temperature=300, dt=1.0, seed=5, precision="f32", steps_per_launch=100)


def run(cfg, M, nsteps=200, x0=None):
    spec = spec_from(cfg, n_walkers=M)
    eng = Engine(spec.to_struct())
    if x0 is None:
        rng = np.random.RandomState(1)
        x0 = np.stack([rng.uniform(18, 28, M), rng.uniform(12, 32, M)])
    eng.set_state(x0)
    eng.begin()
    eng.run(nsteps)
    return eng


def test_split_dump_launch_equals_one_launch(monkeypatch):
    """Walkers [0, 512) record frames through the frame variant on a second stream, the other 1536 run the plain
    kernel: states and frames are those of the single launch."""
    M = 2048
    cfg = dict(C2, dump=True, dump_freq=10, stat_freq=10, dump_walkers=512, dump_capacity=64)
    a = run(cfg, M)
    assert "frames" in a.kernel_name and "f32x2" in a.kernel_name, a.kernel_name
    fa, sa = a.drain_frames(), a.get_state()
    monkeypatch.setenv("NDS_NO_SPLIT", "1")
    b = run(cfg, M)
    fb, sb = b.drain_frames(), b.get_state()
    assert fa.shape == fb.shape == (21, a.frame_width, 512)
    assert np.array_equal(fa, fb)
    for k in ("x", "v", "pe"):
        assert np.array_equal(sa[k], sb[k]), k
    assert a.launch_count > b.launch_count  # two launches per block of steps instead of one
    assert a.frames_f32


def test_packed_frames_match_scalar_frames(monkeypatch):
    M = 1024
    cfg = dict(C2, dump=True, dump_freq=10, stat_freq=10, dump_walkers=M, dump_capacity=64)
    a = run(cfg, M, nsteps=100)
    fa = a.drain_frames()
    names = a.frame_fields()
    monkeypatch.setenv("NDS_NO_PACKED", "1")
    b = run(cfg, M, nsteps=100)
    assert "f32x2" not in b.kernel_name and "f32x2" in a.kernel_name
    fb = b.drain_frames()
    assert fa.shape == fb.shape == (11, 16, M)
    assert np.array_equal(fa[:, names.index("step")], fb[:, names.index("step")])
    assert list(fa[:, names.index("step"), 0].astype(int)) == list(range(0, 101, 10))
    for f in ("x0", "x1", "c0", "pe", "T", "totale"):
        d = np.abs(fa[:, names.index(f)] - fb[:, names.index(f)])
        assert d.max() < 2e-3 * np.abs(fb[:, names.index(f)]).max(), f  # fp32 trajectories, 100 chaotic steps
    # rows are consistent inside a frame: totale = pe + ke (+ biase = 0), T from ke
    i = names.index
    assert np.allclose(fa[1:, i("totale")], fa[1:, i("pe")] + fa[1:, i("ke")], rtol=2e-6)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_inkernel_histogram_counts_every_walker_at_dump_cadence(prec):
    """hist_n: the kernel bins the colvars of ALL walkers at every dump step; equals a numpy histogram of the frames."""
    M = 4096
    cfg = dict(C2, precision=prec, dump=True, dump_freq=10, stat_freq=10, dump_walkers=M, dump_capacity=64,
               hist_n=[48, 40], hist_lo=[0.0, 0.0], hist_hi=[48.0, 40.0])
    eng = run(cfg, M, nsteps=300)
    if prec == "f32":
        assert "frames+hist" in eng.kernel_name, eng.kernel_name
    H = eng.get_histogram()
    fr = eng.drain_frames()
    names = eng.frame_fields()
    steps = fr[:, names.index("step"), 0].astype(int)
    sel = (steps % 10 == 0) & (steps > 0)
    c0 = fr[sel, names.index("c0")].astype(np.float32).ravel()
    c1 = fr[sel, names.index("c1")].astype(np.float32).ravel()
    u = (c0 - np.float32(0.0)) * np.float32(1.0)
    v = (c1 - np.float32(0.0)) * np.float32(1.0)
    ok = (u >= 0) & (v >= 0) & (u < 48) & (v < 40)
    want = np.zeros((40, 48), dtype=np.int64)
    np.add.at(want, (v[ok].astype(int), u[ok].astype(int)), 1)
    assert H.sum() == want.sum() and H.sum() > 0.99 * 30 * M
    assert np.array_equal(H, want)
    # clear on read
    eng.get_histogram(clear=True)
    assert eng.get_histogram().sum() == 0


def test_histogram_without_frames_and_split(monkeypatch):
    """C2 as SURVEY 8d runs it: frames for a walker prefix + histogram of everybody; the walkers beyond the prefix
    run the histogram-only packed variant.  Same counts as the unsplit run."""
    M = 8192
    cfg = dict(C2, dump=True, dump_freq=10, stat_freq=10, dump_walkers=1024, dump_capacity=64,
               hist_n=[48, 40], hist_lo=[0.0, 0.0], hist_hi=[48.0, 40.0])
    a = run(cfg, M, nsteps=200)
    Ha = a.get_histogram()
    monkeypatch.setenv("NDS_NO_SPLIT", "1")
    b = run(cfg, M, nsteps=200)
    assert np.array_equal(Ha, b.get_histogram())
    assert Ha.sum() > 0.99 * 20 * M
    assert np.array_equal(a.drain_frames(), b.drain_frames())
