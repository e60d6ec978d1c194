"""Grid-accumulated metadynamics bias (``mtd_grid``): the footprint of every new hill is added to a grid of
(V, h0 V_0, h1 V_1, h0 h1 V_01) -- the reference's own incremental trick for its FES plot
(``bias/gaus.py:380-393``) -- and the step kernel interpolates (bicubic Hermite).  The exact hill sum
(``bias/gaus.py:236-261``) stays the parity path; these tests bound the distance between the two and check the
multi-rank exchanges of the grid increments."""
import threading

import numpy as np
import pytest

from ndsimulator_b200 import _abi
from ndsimulator_b200.engine import Engine
from tests.helpers import scaled_err, spec_from

pytestmark = pytest.mark.gpu

C5 = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
          x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0, temperature=300, biases=["mtd"], mtd_w=0.05,
          mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, mtd_shared=True, steps=1000)
GRID = dict(mtd_grid=True, mtd_grid_lo=[-8.0, -8.0], mtd_grid_hi=[48.0, 48.0])  # spacing sigma / 4 = 0.5: 113 x 113
MTD2D = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], integrate="langevin", gamma=0.1, dt=1.0,
             temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0,
             mtd_shared=True, steps=2000)


def hermite(t):
    u = 1 - t
    return (np.array([(1 + 2 * t) * u * u, t * t * (3 - 2 * t)]), np.array([t * u * u, -t * t * u]),
            np.array([-6 * t * u, 6 * t * u]), np.array([u * (1 - 3 * t), t * (3 * t - 2)]))


def interp_numpy(G, lo, h, c):
    """Bicubic Hermite interpolation of the engine's grid layout at points c[2][n]: V and g = -dV/dc."""
    ny, nx, _ = G.shape
    V, g0, g1 = np.zeros(c.shape[1]), np.zeros(c.shape[1]), np.zeros(c.shape[1])
    for n in range(c.shape[1]):
        u, w = (c[0, n] - lo[0]) / h[0], (c[1, n] - lo[1]) / h[1]
        i, j = min(int(u), nx - 2), min(int(w), ny - 2)
        a, b, da, db = hermite(u - i)
        as_, bs, das, dbs = hermite(w - j)
        for p in range(2):
            for q in range(2):
                v, vx, vy, vxy = G[j + q, i + p]
                V[n] += a[p] * as_[q] * v + b[p] * as_[q] * vx + a[p] * bs[q] * vy + b[p] * bs[q] * vxy
                g0[n] -= (da[p] * as_[q] * v + db[p] * as_[q] * vx + da[p] * bs[q] * vy + db[p] * bs[q] * vxy) / h[0]
                g1[n] -= (a[p] * das[q] * v + b[p] * das[q] * vx + a[p] * dbs[q] * vy + b[p] * dbs[q] * vxy) / h[1]
    return V, g0, g1


def random_table(n, seed=3):
    rng = np.random.RandomState(seed)
    return rng.uniform(6, 34, size=(n, 2)), rng.uniform(0.005, 0.05, size=n)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_grid_bias_matches_exact_hill_sum(prec):
    """Same 4000 hills, exact sum vs grid + interpolation at 2000 random points.  Bounds written here are what
    DESIGN.md quotes: energy within 1e-4 of the largest bias, force within 2e-3 of the largest force (spacing
    sigma / 4); fp32 adds its rounding."""
    centers, heights = random_table(4000)
    M = 64
    exact = Engine(spec_from(MTD2D, n_walkers=M, precision="f64", mtd_capacity=5000).to_struct())
    grid = Engine(spec_from(MTD2D, n_walkers=M, precision=prec, mtd_capacity=5000, mtd_grid=True,
                            mtd_grid_lo=[-6.0, -6.0], mtd_grid_hi=[46.0, 46.0]).to_struct())
    assert "mtd-grid" in grid.kernel_name, grid.kernel_name
    for e in (exact, grid):
        e.set_hills(centers, heights)
    rng = np.random.RandomState(0)
    pts = np.stack([rng.uniform(2, 38, 2000), rng.uniform(2, 38, 2000)])
    a, b = exact.eval_at(pts), grid.eval_at(pts)
    ev = np.max(np.abs(a["Vb"] - b["Vb"])) / np.max(np.abs(a["Vb"]))
    ef = np.max(np.abs(a["Fb"] - b["Fb"])) / np.max(np.abs(a["Fb"]))
    print(f"mtd_grid {prec}: max |dV| / max V = {ev:.2e}, max |dF| / max F = {ef:.2e}")
    assert ev < (1e-4 if prec == "f64" else 2e-4)
    assert ef < 2e-3
    assert grid.mtd_grid_outside() == 0
    # the FES estimate of the reference (sum of hills on its plot grid) from the accumulated grid
    fa, fb = exact.fes_grid(0.0, 1.0, 40, 0.0, 1.0, 40), grid.fes_grid(0.0, 1.0, 40, 0.0, 1.0, 40)
    assert np.max(np.abs(fa - fb)) < 2e-4 * np.max(fa)


def test_grid_interpolation_is_the_documented_formula():
    """The device interpolation equals a numpy restatement of the bicubic Hermite formula on the downloaded grid."""
    centers, heights = random_table(500, seed=9)
    cfg = dict(MTD2D, mtd_grid=True, mtd_grid_lo=[0.0, 0.0], mtd_grid_hi=[40.0, 40.0], mtd_grid_n=[81, 81])
    eng = Engine(spec_from(cfg, n_walkers=32, precision="f64", mtd_capacity=1000).to_struct())
    eng.set_hills(centers, heights)
    G = eng.get_mtd_grid()
    assert G.shape == (81, 81, 4)
    rng = np.random.RandomState(1)
    pts = np.stack([rng.uniform(1, 39, 300), rng.uniform(1, 39, 300)])
    out = eng.eval_at(pts)
    V, g0, g1 = interp_numpy(G, [0.0, 0.0], [0.5, 0.5], pts)
    assert np.max(np.abs(out["Vb"] - V)) < 1e-13 * max(1.0, np.max(np.abs(V)))
    assert np.max(np.abs(out["Fb"][0] - g0)) < 1e-12 and np.max(np.abs(out["Fb"][1] - g1)) < 1e-12
    # grid values themselves: node (j, i) holds the exact hill sum and its derivatives
    j, i = 40, 44
    X, Y = 0.5 * i, 0.5 * j
    d0, d1 = X - centers[:, 0], Y - centers[:, 1]
    e = heights * np.exp(-(d0 ** 2 + d1 ** 2) / 8.0)
    want = np.array([e.sum(), -(e * d0 / 4.0).sum() * 0.5, -(e * d1 / 4.0).sum() * 0.5,
                     (e * d0 * d1 / 16.0).sum() * 0.25])
    assert np.allclose(G[j, i], want, rtol=1e-10, atol=1e-14)
    # a point outside the box sees no bias and is counted
    far = eng.eval_at(np.array([[45.0], [20.0]]))
    assert far["Vb"][0] == 0.0 and np.all(far["Fb"] == 0.0)


def test_grid_accumulates_like_a_rebuild_and_deposits_like_the_exact_path():
    """Run with depositions: (1) the incrementally accumulated grid equals a grid rebuilt from the final hill table;
    (2) with the same external noise the first well-tempered heights follow the exact engine (the stale-VR rule,
    SURVEY A3, reads the bias from the grid)."""
    M, nsteps = 8, 1000
    rng = np.random.RandomState(5)
    noise = rng.normal(size=(nsteps, 2, M))
    x0 = np.array([[22.0], [24.0]]) + rng.normal(scale=1.0, size=(2, M))
    cfgs = dict(exact=dict(MTD2D), grid=dict(MTD2D, mtd_grid=True, mtd_grid_lo=[-10.0, -10.0], mtd_grid_hi=[58.0, 58.0]))
    eng = {}
    for k, cfg in cfgs.items():
        spec = spec_from(cfg, n_walkers=M, precision="f64", steps_per_launch=100, noise_mode=_abi.NDS_NOISE_EXTERNAL)
        eng[k] = Engine(spec.to_struct())
        eng[k].set_state(x0, np.zeros((2, M)))
        eng[k].set_noise(noise)
        eng[k].begin()
        eng[k].run(nsteps)
    (ce, he), (cg, hg) = eng["exact"].get_hills(), eng["grid"].get_hills()
    assert len(he) == len(hg) == 10 * M
    # trajectories separate only through the interpolation error of the bias force
    assert np.max(np.abs(ce - cg)) < 5e-3 and np.max(np.abs(he - hg)) < 1e-4 * 0.05
    G = eng["grid"].get_mtd_grid()
    fresh = Engine(spec_from(cfgs["grid"], n_walkers=M, precision="f64").to_struct())
    fresh.set_hills(cg, hg)
    G2 = fresh.get_mtd_grid()
    assert np.max(np.abs(G - G2)) < 1e-12 * np.max(np.abs(G2))
    st_e, st_g = eng["exact"].get_state(), eng["grid"].get_state()
    assert np.max(np.abs(st_e["x"] - st_g["x"])) < 5e-3
    assert np.max(np.abs(st_e["biase"] - st_g["biase"])) < 1e-4


def test_grid_in_shared_memory_equals_grid_in_global_memory(monkeypatch):
    """md_block_wide with the TMA-loaded shared-memory copy of the grid and the same kernel reading the grid through
    L2 execute the same arithmetic: identical states (fp32, Philox noise, several depositions)."""
    M = 512
    cfg = dict(C5, **GRID, seed=11)
    rng = np.random.RandomState(2)
    x0 = np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M), rng.uniform(-3, 3, M),
                   np.full(M, 12.0)])
    out = {}
    monkeypatch.setenv("NDS_NO_PAIR", "1")  # the thread-per-walker kernel (small ensembles default to md_pair_grid)
    for mode in ("smem", "global"):
        if mode == "global":
            monkeypatch.setenv("NDS_GRID_NO_SMEM", "1")
        eng = Engine(spec_from(cfg, n_walkers=M, precision="f32", steps_per_launch=100).to_struct())
        assert eng.kernel_name.endswith("md_block_wide<f32,nd5,Mueller5d,Proj5dto2d,langevin,mtd-grid>")
        eng.set_state(x0)
        eng.begin()
        eng.run(500)
        out[mode] = (eng.get_state(), eng.get_mtd_grid(), eng.mtd_grid_outside())
    for k in ("x", "v", "biase", "fixf"):
        assert np.array_equal(out["smem"][0][k], out["global"][0][k]), k
    assert np.array_equal(out["smem"][1], out["global"][1])
    assert out["smem"][2] == out["global"][2] == 0
    assert out["smem"][0]["biase"].max() > 0.05


def c5_start(M, seed=2):
    rng = np.random.RandomState(seed)
    return np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M), rng.uniform(-3, 3, M),
                     np.full(M, 12.0)])


def test_pair_kernel_follows_the_thread_per_walker_kernel(monkeypatch):
    """md_pair_grid (two lanes per walker, nds_pair_kernel.cuh) against md_block_wide: same Philox stream, same
    physics, the force summed in a different order -- the states agree to fp32 rounding after 20 steps and stay close
    over 5 depositions (Langevin dynamics contract); hills, VR and the grid follow.  An odd walker count leaves a
    shadow pair in the last warp."""
    M = 201
    cfg = dict(C5, **GRID, seed=11)
    x0 = c5_start(M)
    out = {}
    for mode in ("pair", "wide"):
        if mode == "wide":
            monkeypatch.setenv("NDS_NO_PAIR", "1")
        eng = Engine(spec_from(cfg, n_walkers=M, precision="f32", steps_per_launch=100).to_struct())
        assert eng.kernel_name.startswith("md_pair_grid|md_block_wide<f32,nd5,Mueller5d")
        eng.set_state(x0)
        eng.begin()
        eng.run(20)
        s20 = eng.get_state()
        eng.run(480)
        out[mode] = (s20, eng.get_state(), eng.get_hills(), eng.get_mtd_grid(), eng.mtd_grid_outside())
    a, b = out["pair"], out["wide"]
    d20 = {k: np.max(np.abs(a[0][k] - b[0][k])) for k in ("x", "v", "f", "fixf", "pe", "biase")}
    d500 = {k: np.max(np.abs(a[1][k] - b[1][k])) for k in ("x", "v", "biase")}
    print("pair vs wide after 20 steps:", d20, "after 500:", d500)
    assert d20["x"] < 1e-4 and d20["v"] < 1e-5 and d20["pe"] < 1e-4
    assert not np.array_equal(a[0]["x"], x0) and np.max(np.abs(a[0]["x"] - x0)) > 0.05  # the walkers did move
    assert d500["x"] < 5e-2 and np.median(np.abs(a[1]["x"] - b[1]["x"])) < 1e-3
    (ca, ha), (cb, hb) = a[2], b[2]
    assert len(ha) == len(hb) == 5 * M
    assert np.max(np.abs(ca - cb)) < 5e-2 and np.max(np.abs(ha - hb)) < 1e-4 * 0.05
    assert np.max(np.abs(a[3] - b[3])) < 2e-3 * np.max(np.abs(b[3]))
    assert a[4] == b[4] == 0
    assert a[1]["biase"].max() > 0.05


def test_pair_kernel_launch_boundaries_and_fp64():
    """(1) Launch boundaries at odd and even steps do not change the pair kernel's results (its launch entry
    evaluates the force with the loop's own arithmetic; steps are drawn in (even, odd) pairs).  (2) Against the fp64
    engine under the same Philox seed the walkers stay within fp32 distance over 60 steps."""
    M = 64
    cfg = dict(C5, **GRID, seed=5, mtd_dep_freq=1000)  # no deposition inside the window
    x0 = c5_start(M, seed=7)

    def run(prec, chunks):
        eng = Engine(spec_from(cfg, n_walkers=M, precision=prec, steps_per_launch=1000).to_struct())
        eng.set_state(x0)
        eng.begin()
        for c in chunks:
            eng.run(c)
        return eng.get_state()

    one = run("f32", [60])
    for chunks in ([7, 13, 40], [1, 1, 58], [30, 30], [2, 3, 5, 6, 44], [59, 1]):
        st = run("f32", chunks)
        for k in ("x", "v", "f", "pe"):
            assert np.array_equal(one[k], st[k]), (chunks, k)
    ref = run("f64", [60])
    # the fp64 engine extends every 21-bit field by a second Philox call: the two noise streams differ by the cell
    # width (2^-21 relative in the radius variate), like the fp32 / fp64 pairs of tests/test_fuzz_gpu.py
    assert np.max(np.abs(one["x"] - ref["x"])) < 1e-2
    assert scaled_err(one["pe"], ref["pe"]) < 5e-3


def test_grid_one_dimensional_colvar():
    """colvardim 1: (V, h V') per node, cubic Hermite along one axis."""
    cfg = dict(ndim=2, mass=1.0, potential="Mueller2d", colvar_name="XmY", x0=[22.0, 24.0], integrate="langevin",
               gamma=0.1, dt=1.0, temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=1.5, mtd_dep_freq=10,
               mtd_biasf=10.0, mtd_shared=True, steps=100)
    rng = np.random.RandomState(12)
    centers, heights = rng.uniform(-10, 10, size=(700, 1)), rng.uniform(0.01, 0.05, 700)
    exact = Engine(spec_from(cfg, n_walkers=16, precision="f64", mtd_capacity=1000).to_struct())
    grid = Engine(spec_from(cfg, n_walkers=16, precision="f64", mtd_capacity=1000, mtd_grid=True,
                            mtd_grid_lo=[-20.0], mtd_grid_hi=[20.0]).to_struct())
    for e in (exact, grid):
        e.set_hills(centers, heights)
    assert grid.get_mtd_grid().shape == (1, 108, 4)
    pts = np.stack([rng.uniform(10, 30, 400), rng.uniform(10, 30, 400)])
    pts = pts[:, np.abs(pts[0] - pts[1]) < 15]
    a, b = exact.eval_at(pts), grid.eval_at(pts)
    assert scaled_err(b["Vb"], a["Vb"]) < 1e-4 and scaled_err(b["Fb"], a["Fb"]) < 2e-3


def _two_rank_grid_run(attach, M=64, nsteps=400, prec="f32"):
    cfg = dict(C5, **GRID, seed=21, precision=prec, steps_per_launch=100)
    rng = np.random.RandomState(8)
    x0 = np.stack([rng.uniform(8, 30, 2 * M), rng.uniform(-3, 3, 2 * M), rng.uniform(8, 30, 2 * M),
                   rng.uniform(-3, 3, 2 * M), np.full(2 * M, 12.0)])
    one = Engine(spec_from(cfg, n_walkers=2 * M).to_struct())
    one.set_state(x0)
    one.begin()
    one.run(nsteps)
    engines = [Engine(spec_from(cfg, n_walkers=M, walker_offset=r * M, n_walkers_global=2 * M).to_struct())
               for r in range(2)]
    barrier = threading.Barrier(2)
    attach(engines, barrier)
    errors = []

    def drive(r):
        try:
            e = engines[r]
            e.set_state(x0[:, r * M:(r + 1) * M])
            e.begin()
            e.run(nsteps)
            e.sync()
        except Exception as ex:  # pragma: no cover
            errors.append(ex)
            barrier.abort()

    threads = [threading.Thread(target=drive, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    return one, engines


def _check_two_ranks(one, engines, M):
    G = one.get_mtd_grid()
    g0, g1 = engines[0].get_mtd_grid(), engines[1].get_mtd_grid()
    assert np.array_equal(g0, g1)                       # rank-ordered sums: bit-identical grids on every rank
    assert np.max(np.abs(g0 - G)) < 2e-5 * np.max(np.abs(G))   # one engine sums all hills in one order instead
    ref = one.get_state()
    for r in range(2):
        st = engines[r].get_state()
        assert np.max(np.abs(st["x"] - ref["x"][:, r * M:(r + 1) * M])) < 5e-2
        assert engines[r].n_hills() == 4 * M            # the table keeps this rank's own hills only
    c = np.concatenate([engines[r].get_hills()[0] for r in range(2)])
    assert c.shape == (8 * M, 2)


def test_two_rank_grid_allreduce_emulated_on_one_gpu():
    """N > 1 path with the all-reduce exchange of the grid increment (the role of ncclAllReduce): two engines
    driven by two host threads; the callback sums both increment buffers in rank order."""
    import torch
    from ndsimulator_b200.distributed import _DevSegment
    M = 64
    bufs = [None, None]

    def attach(engines, barrier):
        def make_cb(r):
            def cb(ptr, nbytes, off):
                assert off == -1
                bufs[r] = torch.as_tensor(_DevSegment(ptr, nbytes), device="cuda").view(torch.float32)
                barrier.wait()
                total = bufs[0] + bufs[1]
                torch.cuda.synchronize()
                barrier.wait()
                bufs[r].copy_(total)
                torch.cuda.synchronize()
                barrier.wait()
                return 0
            return cb
        for r in range(2):
            engines[r].set_exchange(make_cb(r))

    one, engines = _two_rank_grid_run(attach, M=M)
    _check_two_ranks(one, engines, M)
    assert engines[0].exchange_info()["n_exchanges"] == 4


def test_two_rank_grid_peer_slots_emulated_on_one_gpu():
    """Fused accumulate + exchange over peer pointers: every engine writes its grid increment into its slot of BOTH
    windows (nds_set_peer_tables), the callback is only a host barrier (ranks share one GPU here, so the device
    barrier stays off), then each engine adds the slots in rank order."""
    M = 64

    def attach(engines, barrier):
        ptrs = [e.device_pointers()["hills"] for e in engines]

        def cb(ptr, nbytes, off):
            assert nbytes == 0
            barrier.wait()
            return 0
        for r in range(2):
            engines[r].set_peer_tables(ptrs, r)
            engines[r].set_exchange(cb)

    one, engines = _two_rank_grid_run(attach, M=M)
    _check_two_ranks(one, engines, M)


def test_grid_checkpoint_resume_is_exact():
    M = 128
    cfg = dict(C5, **GRID, seed=4, precision="f32", steps_per_launch=100)
    rng = np.random.RandomState(3)
    x0 = np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M), rng.uniform(-3, 3, M),
                   np.full(M, 12.0)])
    a = Engine(spec_from(cfg, n_walkers=M).to_struct())
    a.set_state(x0)
    a.begin()
    a.run(350)
    ck = a.checkpoint()
    a.run(450)
    b = Engine(spec_from(cfg, n_walkers=M).to_struct())
    b.restore(ck)
    b.run(450)
    sa, sb = a.get_state(), b.get_state()
    for k in ("x", "v", "biase"):
        assert np.array_equal(sa[k], sb[k]), k
    assert np.array_equal(a.get_mtd_grid(), b.get_mtd_grid())


def test_grid_mode_free_energy_surface_follows_the_exact_mode():
    """The same multiple-walker run (same Philox noise) with the exact hill sum and with the grid-accumulated bias:
    Langevin dynamics contract, so the two ensembles stay close and the free-energy estimates built from their hill
    tables (sum of hills on the reference's 48 x 40 grid) agree far inside the north star's 0.1 kT RMS."""
    from ndsimulator_b200.constant import kB
    M, n = 256, 3000
    cfg = dict(C5, mtd_w=0.002, steps=n, seed=8)
    rng = np.random.RandomState(0)
    x0 = np.stack([rng.uniform(14, 26, M), rng.uniform(-2, 2, M), rng.uniform(20, 32, M), rng.uniform(-2, 2, M),
                   np.full(M, 12.0)])
    fes = []
    for extra in ({}, GRID):
        eng = Engine(spec_from(dict(cfg, **extra), n_walkers=M, precision="f32", steps_per_launch=100).to_struct())
        eng.set_state(x0)
        eng.begin()
        eng.run(n)
        assert eng.n_hills() == 30 * M
        fes.append(eng.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40))
        if extra:
            assert "mtd-grid" in eng.kernel_name and eng.mtd_grid_outside() == 0
    kT = kB * 300
    mask = fes[0] > 2 * kT
    assert mask.sum() > 100
    d = (fes[1] - fes[0])[mask]
    rms = np.sqrt(((d - d.mean()) ** 2).mean())
    assert rms < 0.05 * kT, rms / kT
