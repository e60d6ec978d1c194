"""CPU: pins the C oracle (oracle/nds_oracle.c) against fixtures generated from the UNMODIFIED Python
reference (oracle/make_golden.py -> tests/golden/*.json).  Tolerances: 1e-12 relative for point
evaluations (libm vs NumPy SIMD exp differ by <= 1 ulp), 1e-9 for 10^4-step NVE trajectories (ulp
differences amplified by the dynamics); stochastic trajectories must match too because the oracle
reproduces numpy.random.RandomState (MT19937 + legacy polar Gaussian) bit for bit."""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import load_golden, spec_from, scaled_err


def test_rng_matches_numpy_randomstate():
    for seed in (0, 7, 12345, 2**32 - 1):
        rs = np.random.RandomState(seed)
        ref = np.concatenate([rs.normal(0, 1, 1), rs.normal(0, 1, 5), rs.normal(0, 1, 2)])
        assert np.array_equal(oracle.rng_normals(seed, 8), ref)
        assert np.array_equal(oracle.rng_uniforms(seed, 9), np.random.RandomState(seed).rand(9))


def test_survey_spot_values():
    """SURVEY.md section 8c spot values of Mueller2d.compute (shipped code, fp64)."""
    spec = spec_from(dict(ndim=2, potential="Mueller2d", x0=[22.0, 24.0], mass=1.0, integrate="nve"))
    o = oracle.OracleSim(spec.to_struct(), seeds=[0])
    pts = np.array([[22, 24], [21, 17], [12, 12], [18, 18], [22, 30]], dtype=float).T
    r = o.eval_at(pts)
    V = [-0.4263196475207508, -0.4042981723685126, 0.07779167017139693, -0.35559755602211907,
         -1.257182522254452]
    F = [[-0.1272979074308674, 0.13497268481028946], [0.030788854319483673, -0.022756667565805022],
         [0.017003641640205507, 0.0677566365004881], [-0.006299322695502431, 0.017506570016558233],
         [0.014274079368294798, 0.013948684867191781]]
    assert np.allclose(r["V"], V, rtol=1e-13, atol=0)
    assert np.allclose(r["F"].T, F, rtol=1e-12, atol=0)


def test_constants():
    g = load_golden("constants.json")
    spec = spec_from(dict(ndim=2, potential="Mueller2d", x0=[22.0, 24.0], mass=1.0, gamma=0.1,
                          temperature=300, dt=1.0))
    c = oracle.OracleSim(spec.to_struct(), seeds=[0]).constants()
    for k in ("m", "kBT", "c1", "c2"):
        assert c[k] == pytest.approx(g["C2_langevin_gamma0.1"][k], rel=1e-15)
    # SURVEY 8a4 probe values
    assert c["c1"] == pytest.approx(0.951229424500714, rel=1e-15)
    assert c["c2"] == pytest.approx(0.004858219817722203, rel=1e-14)
    spec = spec_from(dict(ndim=5, potential="Mueller5d", colvar_name="Proj5dto2d",
                          true_colvar_name="Proj5dto2d", x0=[18.0, 0, 18.0, 0, 100.0], mass=1.0,
                          integrate="2nd-langevin", md_gamma=0.01, dt=1.0, temperature=300.0))
    c = oracle.OracleSim(spec.to_struct(), seeds=[0]).constants()
    for k in ("c1", "c2", "c3", "c4", "c5"):
        assert c[k] == pytest.approx(g["C4_5d_umb"][k], rel=1e-14)


@pytest.mark.parametrize("idx", range(23))
def test_eval_points(idx):
    rec = load_golden("eval_points.json")[idx]
    cfg = dict(rec["config"], x0=rec["points"][0], mass=1.0, integrate="nve")
    spec = spec_from(cfg)
    o = oracle.OracleSim(spec.to_struct(), seeds=[0])
    pts = np.array(rec["points"]).T
    r = o.eval_at(pts)
    n = pts.shape[1]
    assert np.allclose(r["V"], rec["V"], rtol=1e-12, atol=1e-300)
    assert np.allclose(r["F"].T, np.array(rec["F"]), rtol=1e-11, atol=1e-18)
    assert np.allclose(r["colv"].T, np.array(rec["colv"]), rtol=1e-14, atol=0)
    J = np.array(rec["J"]).reshape(n, spec.colvardim, spec.ndim)
    assert np.allclose(np.moveaxis(r["J"], 2, 0), J, rtol=1e-14, atol=0)


@pytest.mark.parametrize("idx", range(10))
def test_bias_points(idx):
    rec = load_golden("bias_points.json")[idx]
    cfg = dict(rec["config"], mass=1.0, integrate="nve")
    spec = spec_from(cfg)
    o = oracle.OracleSim(spec.to_struct(), seeds=[0])
    if "hills_centers" in rec and "mtd" in cfg["biases"]:
        o.set_hills(np.array(rec["hills_centers"]), np.array(rec["hills_heights"]))
    r = o.eval_at(np.array(rec["points"]).T)
    assert np.allclose(r["Vb"], rec["Vb"], rtol=1e-12, atol=1e-300)
    assert np.allclose(r["Fb"].T, np.array(rec["Fb"]), rtol=1e-11, atol=1e-16)


TRAJ = load_golden("trajectories.json")


@pytest.mark.parametrize("name", sorted(TRAJ))
def test_trajectory(name):
    rec = TRAJ[name]
    cfg = rec["config"]
    spec = spec_from(cfg)
    o = oracle.OracleSim(spec.to_struct(), seeds=[cfg["seed"]])
    o.set_state(spec.x0.reshape(-1, 1))
    o.begin()
    st = o.get_state()
    assert np.allclose(st["v"][:, 0], rec["v0"], rtol=1e-15, atol=0)  # Maxwell-Boltzmann draw
    for k in ("pe", "ke", "T", "totale"):
        assert st[k][0] == pytest.approx(rec["begin"][k], rel=1e-13)
    every = rec["sample_every"]
    traj = o.run(rec["steps"][-1], traj_every=every)
    fr = o.split_frame(traj)
    tol = 1e-9
    for key in ("x", "v", "f", "fixf", "colv"):
        ref = np.array(rec[key])
        got = fr[key][:, :, 0]
        assert scaled_err(got, ref) < tol, key
    for key in ("pe", "ke", "biase", "totale", "T"):
        assert scaled_err(fr[key][:, 0], np.array(rec[key])) < tol, key
    assert o.time == pytest.approx(rec["final_time"], rel=0, abs=0)
    if "hills_heights" in rec:
        c, w = o.get_hills()
        assert np.allclose(w, rec["hills_heights"], rtol=1e-10)
        assert np.allclose(c, np.array(rec["hills_centers"]), rtol=1e-10)
        assert o.get_vr()[0] == pytest.approx(rec["VR"], rel=1e-9)
    if "hills_sigma2" in rec:  # adapsig_geo: Gaussian._sigma_history, J.J^T (first hill) / J.J^T sigma^2
        assert np.allclose(o.get_hill_sigma2(), rec["hills_sigma2"], rtol=1e-12)


def test_mtd_first_hill_heights():
    """first hill heights of examples/2d-mtd.yaml, seed 0."""
    w = np.array(TRAJ["2d_mtd_seed0"]["hills_heights"][:6])
    assert np.allclose(w, [0.05, 0.04033116, 0.03563761, 0.03074257, 0.03092146, 0.03591853], atol=5e-9)


@pytest.mark.parametrize("which", ["shots", "diffusion"])
def test_committor_shots(which):
    """24 shots of examples/2d-committor.yaml; 8 shots started inside basin 0 with diffusion_time = 37
    (committor.py:100-101: the first tested update is the one made at NDRun.step 37, so they exit at step 38)."""
    g = load_golden("committor.json")
    if which == "diffusion":
        g = g["diffusion"]
        assert all(s["step"] == 38 and s["basin"] == 0 for s in g["shots"])
    cfg = g["config"]
    shots = g["shots"]
    spec = spec_from(cfg, n_walkers=len(shots))
    o = oracle.OracleSim(spec.to_struct(), seeds=[s["seed"] for s in shots])
    o.set_state(np.repeat(spec.x0.reshape(-1, 1), len(shots), axis=1))
    o.begin()
    assert np.allclose(o.get_state()["v"].T, np.array([s["v0"] for s in shots]), rtol=1e-15)
    o.set_threads(oracle.max_threads())
    o.run(cfg["steps"])
    basin, exit_step = o.get_commit()
    assert list(basin) == [s["basin"] for s in shots]
    assert list(exit_step) == [s["step"] for s in shots]
    assert np.allclose(o.get_state()["x"].T, np.array([s["x"] for s in shots]), rtol=1e-9)


def test_committor_temperature_zero_minimize():
    """T = 0 committor: the inner engine is the fixed-step steepest descent of engines/minimize.py;
    deterministic, so basin, exit step, final position and energy must equal the reference's."""
    g = load_golden("committor_T0.json")
    runs = g["runs"]
    spec = spec_from(g["config"], n_walkers=len(runs))
    o = oracle.OracleSim(spec.to_struct(), seeds=[0] * len(runs))
    o.set_state(np.array([r["x0"] for r in runs]).T)
    o.begin()
    o.run(g["config"]["steps"])
    basin, exit_step = o.get_commit()
    assert list(basin) == [r["basin"] for r in runs]
    assert list(exit_step) == [r["step"] for r in runs]
    st = o.get_state()
    assert np.allclose(st["x"].T, np.array([r["x"] for r in runs]), rtol=1e-12)
    assert np.allclose(st["pe"], [r["pe"] for r in runs], rtol=1e-12)
    assert np.allclose(st["colv"].T, np.array([r["colv"] for r in runs]), rtol=1e-12)


FES = load_golden("fes_projection.json")


def fes_case_grid(rec):
    """(x0, dx, nx[, y0, dy, ny]) of np.arange(b0, b1, inc) per colvar (io/plot.py:483-488)."""
    args = []
    for (b0, b1), inc in zip(rec["boundary"], rec["increment"]):
        args += [b0, inc, len(np.arange(b0, b1, inc))]
    return args


@pytest.mark.parametrize("name", sorted(FES))
def test_fes_projection_matches_reference(name):
    """Gaussian.projection (gaus.py:365-438) as the reference's Plot calls it, incremental calls included: the map
    after call k is the sum of the hills deposited by then."""
    rec = FES[name]
    spec = spec_from(rec["config"])
    centers, heights = np.array(rec["hills_centers"]), np.array(rec["hills_heights"])
    for n, want in zip(rec["hills_at_call"], rec["maps"]):
        o = oracle.OracleSim(spec.to_struct(), seeds=[0])
        o.set_hills(centers[:n], heights[:n])
        if "hills_sigma2" in rec:
            o.set_hill_sigma2(rec["hills_sigma2"][:n])
        g = o.fes_grid(*fes_case_grid(rec))
        want = np.array(want).reshape(rec["shape"])
        assert g.reshape(want.shape) == pytest.approx(want, rel=1e-12, abs=1e-300)
