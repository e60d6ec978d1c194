"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bars (BASELINE.json north_star):
  * deterministic paths (NVE, zero-friction Langevin, bias / force evaluation at given configurations,
    and every integrator when the Gaussian numbers are supplied externally) in fp64:
      - per-step relative error <= 1e-10, checked over 10^4 steps (one-step map along the oracle
        trajectory, every state advanced by one step on the GPU);
      - whole 10^4-step trajectories within 1e-9 of the oracle (ulp-level differences of exp / FMA
        contraction amplified by the dynamics; the oracle itself is within 6e-14 of the NumPy reference);
  * stochastic paths (Philox noise): statistical agreement with oracle ensembles (KS tests, effective
    temperature of the reference's integrator quirk A1/A6, committor fractions) -- tests/test_stats_gpu.py.
"""
import numpy as np
import pytest

from ndsimulator_b200 import _abi
from ndsimulator_b200.engine import Engine
from oracle import oracle
from tests.helpers import relerr, load_golden, spec_from, scaled_err

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-10   # per-step relative error, fp64
TRAJ_TOL = 1e-9    # 10^4-step trajectory, fp64


def make_pair(cfg, M=1, seeds=None, **extra):
    spec = spec_from(cfg, n_walkers=M, **extra)
    st = spec.to_struct()
    eng = Engine(st)
    orc = oracle.OracleSim(st, seeds=seeds if seeds is not None else list(range(M)))
    return spec, eng, orc


# ------------------------------------------------------------------------------ point evaluations
@pytest.mark.parametrize("idx", range(23))
def test_eval_points(idx):
    rec = load_golden("eval_points.json")[idx]
    cfg = dict(rec["config"], x0=rec["points"][0], mass=1.0, integrate="nve")
    spec, eng, orc = make_pair(cfg)
    pts = np.array(rec["points"]).T
    r = eng.eval_at(pts)
    n = pts.shape[1]
    # against the fixtures generated from the Python reference
    assert np.allclose(r["V"], rec["V"], rtol=1e-11, atol=1e-300)
    assert np.allclose(r["F"].T, np.array(rec["F"]), rtol=1e-10, atol=1e-17)
    assert np.allclose(r["colv"].T, np.array(rec["colv"]), rtol=1e-13, atol=0)
    J = np.array(rec["J"]).reshape(n, spec.colvardim, spec.ndim)
    assert np.allclose(np.moveaxis(r["J"], 2, 0), J, rtol=1e-13, atol=0)
    # and against the oracle on more points
    rng = np.random.RandomState(idx)
    lo, hi = pts.min(), pts.max()
    more = rng.uniform(lo, hi, size=(spec.ndim, 2000))
    a, b = eng.eval_at(more), orc.eval_at(more)
    for k in ("V", "F", "colv", "J"):
        assert scaled_err(a[k], b[k]) < 1e-11, k


@pytest.mark.parametrize("idx", range(10))
def test_bias_points(idx):
    rec = load_golden("bias_points.json")[idx]
    cfg = dict(rec["config"], mass=1.0, integrate="nve")
    if "mtd" in cfg["biases"]:
        cfg["mtd_capacity"] = 64
    spec, eng, orc = make_pair(cfg)
    if "hills_centers" in rec and "mtd" in cfg["biases"]:
        eng.set_hills(np.array(rec["hills_centers"]), np.array(rec["hills_heights"]))
        orc.set_hills(np.array(rec["hills_centers"]), np.array(rec["hills_heights"]))
    pts = np.array(rec["points"]).T
    r = eng.eval_at(pts)
    assert np.allclose(r["Vb"], rec["Vb"], rtol=1e-11, atol=1e-300)
    assert np.allclose(r["Fb"].T, np.array(rec["Fb"]), rtol=1e-10, atol=1e-15)
    rng = np.random.RandomState(idx)
    more = rng.uniform(pts.min(), pts.max(), size=(spec.ndim, 1000))
    a, b = eng.eval_at(more), orc.eval_at(more)
    assert scaled_err(a["Vb"], b["Vb"]) < 1e-11
    assert scaled_err(a["Fb"], b["Fb"]) < 1e-11


# -------------------------------------------------------------------- one-step map over 10^4 steps
MD2D = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md", temperature=300,
            dt=1.0, seed=7)


@pytest.mark.parametrize("integrate,gamma", [("nve", 0.0), ("langevin", 0.0)])
def test_one_step_map_10k(integrate, gamma):
    """Every state of a 10^4-step oracle trajectory is advanced by ONE step on the GPU (10^4 walkers in
    one launch) and compared with the oracle's next state: relative error per step <= 1e-10."""
    n = 10000
    cfg = dict(MD2D, integrate=integrate, gamma=gamma)
    spec, _, orc = make_pair(cfg, M=1, seeds=[7])
    orc.set_state(spec.x0.reshape(-1, 1))
    orc.begin()
    v0 = orc.get_state()["v"].copy()
    x0 = spec.x0.reshape(-1, 1).copy()
    traj = orc.run(n, traj_every=1)
    fr = orc.split_frame(traj)
    xs = np.concatenate([x0.T, fr["x"][:-1, :, 0]], axis=0).T  # states 0..n-1
    vs = np.concatenate([v0.T, fr["v"][:-1, :, 0]], axis=0).T
    spec2 = spec_from(cfg, n_walkers=n)
    eng = Engine(spec2.to_struct())
    eng.set_state(xs, vs)
    eng.begin()
    eng.run(1)
    st = eng.get_state()
    assert scaled_err(st["x"], fr["x"][:, :, 0].T) < STEP_TOL
    assert scaled_err(st["v"], fr["v"][:, :, 0].T) < STEP_TOL
    assert scaled_err(st["pe"], fr["pe"][:, 0]) < STEP_TOL
    assert scaled_err(st["f"], fr["f"][:, :, 0].T) < STEP_TOL
    assert scaled_err(st["totale"], fr["totale"][:, 0]) < STEP_TOL
    # per element, not only against the largest entry of the vector: every component of every state within 1e-10 of
    # its own magnitude (absolute floor: 1e-3 of the largest entry, where zero crossings of v and f live)
    for k, ref in (("x", fr["x"][:, :, 0].T), ("v", fr["v"][:, :, 0].T), ("f", fr["f"][:, :, 0].T),
                   ("pe", fr["pe"][:, 0]), ("totale", fr["totale"][:, 0])):
        assert relerr(st[k], ref, floor=1e-3) < STEP_TOL, k


# ------------------------------------------------------------------ golden deterministic trajectories
kB_ = 8.6173303e-5
TRAJ = load_golden("trajectories.json")
DETERMINISTIC = ["2d_nve_seed7", "2d_langevin_gamma0_seed7", "2d_rescale_seed3",
                 "1d_doublewell_nve_seed4", "2d_threehole_nve_seed5", "2d_minimize"]


@pytest.mark.parametrize("name", DETERMINISTIC)
def test_trajectories(name):
    """fp64 trajectories against the fixtures of the Python reference (initial velocity injected)."""
    rec = TRAJ[name]
    cfg = rec["config"]
    spec = spec_from(cfg, steps_per_launch=rec["sample_every"])
    eng = Engine(spec.to_struct())
    eng.set_state(spec.x0.reshape(-1, 1), np.array(rec["v0"]).reshape(-1, 1))
    eng.begin()
    st = eng.get_state()
    for k in ("pe", "ke", "T", "totale"):
        assert st[k][0] == pytest.approx(rec["begin"][k], rel=1e-12)
    got = {k: [] for k in ("x", "v", "f", "colv", "pe", "ke", "totale", "T")}
    prev = 0
    for s in rec["steps"]:
        eng.run(s - prev)
        prev = s
        st = eng.get_state()
        for k in got:
            got[k].append(st[k][..., 0].copy())
    for k in got:
        assert scaled_err(np.array(got[k]).reshape(len(rec["steps"]), -1),
                          np.array(rec[k]).reshape(len(rec["steps"]), -1)) < TRAJ_TOL, k
    assert eng.step == rec["steps"][-1]
    assert eng.time == rec["final_time"]


# ------------------------------------------ the reference's STOCHASTIC runs, given the reference's own Gaussians
STOCHASTIC = ["2d_langevin_gamma0.1_seed0", "2d_md_yaml_as_shipped_seed0", "2d_langevin2_seed3", "5d_umb_seed0",
              "2d_umb_seed1", "2d_mtd_seed0", "5d_mtd_seed2", "1d_doublewell_langevin_seed4",
              "2d_doublewell_pe_seed6", "2d_mtd_adapsig_geo_XmY_seed2", "5d_mtd_adapsig_geo_Proj5dto1d_seed3",
              "1d_mtd_adapsig_hist_doublewell_seed5", "1d_mtd_adapsig_hist_doublewell_stat5_seed6"]


def reference_normals(seed, ndim, nsteps, per_step):
    """The numbers RandomState(seed) handed to the reference: ndim separate normal(0, 1, 1) draws for the initial
    velocities (modify.py:44-54), then normal(0, 1, ndim) once (langevin) or twice (2nd-langevin) per step
    (md.py:120,127-128).  numpy's legacy polar method caches every second Gaussian, so the call pattern matters."""
    rs = np.random.RandomState(seed)
    v = np.array([rs.normal(0, 1, 1)[0] for _ in range(ndim)])
    table = np.zeros((nsteps, per_step, 1))
    for k in range(nsteps):
        for j in range(per_step // ndim):
            table[k, j * ndim:(j + 1) * ndim, 0] = rs.normal(0, 1, ndim)
    return v, table


@pytest.mark.parametrize("name", STOCHASTIC)
def test_stochastic_trajectories_with_the_references_own_noise(name):
    """fp64 GPU against the fixtures of the reference's stochastic runs (Langevin 1st / 2nd order, umbrella,
    metadynamics incl. adapsig_geo, PE bias): the Gaussian numbers are rebuilt with numpy exactly as the reference
    drew them and supplied as external noise, so every sampled state must agree within the trajectory tolerance."""
    rec = TRAJ[name]
    cfg = rec["config"]
    spec = spec_from(cfg, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=rec["sample_every"])
    nd = spec.ndim
    per_step = {"langevin": nd, "2nd-langevin": 2 * nd}[spec.integrate]
    nsteps = rec["steps"][-1]
    z0, table = reference_normals(cfg["seed"], nd, nsteps, per_step)
    m = 104.23315626367025 * cfg["mass"]
    t_init = cfg.get("initial_temperature", cfg["temperature"])
    assert np.allclose(z0 * np.sqrt(kB_ * t_init / m), rec["v0"], rtol=1e-13)   # same stream position
    eng = Engine(spec.to_struct())
    eng.set_state(spec.x0.reshape(-1, 1), np.array(rec["v0"]).reshape(-1, 1))
    eng.set_noise(table)
    eng.begin()
    got = {k: [] for k in ("x", "v", "f", "fixf", "colv", "pe", "ke", "biase", "totale", "T")}
    prev = 0
    for s_ in rec["steps"]:
        eng.run(s_ - prev)
        prev = s_
        st = eng.get_state()
        for k in got:
            got[k].append(st[k][..., 0].copy())
    for k in got:
        assert scaled_err(np.array(got[k]).reshape(len(rec["steps"]), -1),
                          np.array(rec[k]).reshape(len(rec["steps"]), -1)) < TRAJ_TOL, k
    assert eng.time == rec["final_time"]
    if "hills_heights" in rec:
        c, h = eng.get_hills()
        assert np.allclose(h, rec["hills_heights"], rtol=1e-9) and np.allclose(c, np.array(rec["hills_centers"]), rtol=1e-9)
        assert eng.get_vr()[0] == pytest.approx(rec["VR"], rel=1e-8)
    if "hills_sigma2" in rec:  # adapsig_geo: J.J^T (* sigma^2); history-based adapsig: weighted variance of Stat.colv
        assert np.allclose(eng.get_hill_sigma2(), rec["hills_sigma2"], rtol=1e-12 if cfg.get("mtd_adapsig_geo") else 1e-9)


# -------------------------------------------------- every integrator / bias with externally given noise
def noise_table(seed, nsteps, nn, M):
    return np.random.RandomState(seed).normal(size=(nsteps, nn, M))


EXT_CASES = {
    "langevin_2d": (dict(MD2D, integrate="langevin", gamma=0.1), 2000, 32),
    "langevin2_2d": (dict(MD2D, integrate="2nd-langevin", gamma=0.01), 2000, 32),
    "langevin_1d_dw": (dict(ndim=1, mass=1.0, potential="DoubleWell1d", potential_K1=1.5, x0=30.0,
                            gamma=0.1, dt=1.0, temperature=300), 2000, 16),
    "umb_5d": (dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
                    true_colvar_name="Proj5dto2d", x0=[18.0, 0.0, 18.0, 0.0, 100.0], biases=["umb"],
                    umb_k=0.1, umb_r0=[18, 18], integrate="2nd-langevin", md_gamma=0.01, dt=1.0,
                    temperature=300.0), 2000, 16),
    "umb2_2d_rotate": (dict(MD2D, integrate="langevin", gamma=0.1, colvar_name="Rotate2d",
                            biases=["umb"], umb_n=2, umb_0_k=0.1, umb_0_r0=[30, 0], umb_1_k=0.05,
                            umb_1_r0=[28, 2]), 1000, 8),
    "pe_doublewell": (dict(ndim=2, mass=1.0, potential="DoubleWell", x0=[20.0, 12.0], gamma=0.05,
                           biases=["pe"], pe_thred=-0.45, dt=1.0, temperature=300), 1500, 8),
    "threehole_5d": (dict(ndim=5, mass=1.0, potential="ThreeHole5d", colvar_name="Proj5dto2d",
                          true_colvar_name="Proj5dto2d", x0=[7.0, 1.0, 9.0, 1.0, 3.0], gamma=0.1,
                          dt=0.5, temperature=3000), 1000, 8),
    "mueller5dshl_v2": (dict(ndim=5, mass=1.0, potential="Mueller5dshl", colvar_name="Proj5dto2dv2",
                             true_colvar_name="Proj5dto2dv2", x0=[18.0, 3.0, 18.0, 1.0, 1.0],
                             gamma=0.1, dt=1.0, temperature=300), 1000, 8),
}


@pytest.mark.parametrize("name", sorted(EXT_CASES))
def test_external_noise_trajectory(name):
    """With the Gaussian numbers supplied by the caller every integrator is deterministic: M walkers,
    each with its own noise, must follow the oracle within the trajectory tolerance."""
    cfg, nsteps, M = EXT_CASES[name]
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=250)
    nn = {"langevin": spec.ndim, "2nd-langevin": 2 * spec.ndim}[spec.integrate]
    noise = noise_table(11, nsteps, nn, M)
    rng = np.random.RandomState(5)
    x0 = spec.x0.reshape(-1, 1) + rng.normal(scale=0.3, size=(spec.ndim, M))
    v0 = rng.normal(scale=0.01, size=(spec.ndim, M))
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        sim.set_noise(noise)
        sim.begin()
        sim.run(nsteps)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "f", "fixf", "colv", "pe", "ke", "biase", "totale", "T"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k


def test_mtd_single_walker_external_noise():
    """examples/2d-mtd.yaml semantics (deposition schedule, stale-VR tempering, force refresh)."""
    cfg = dict(MD2D, x0=[22.0, 30.0], integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05,
               mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, steps=3000)
    spec, eng, orc = make_pair(cfg, M=1, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=1000)
    noise = noise_table(3, 3000, 2, 1)
    for sim in (eng, orc):
        sim.set_state(spec.x0.reshape(-1, 1), np.array([[0.01], [-0.02]]))
        sim.set_noise(noise)
        sim.begin()
        sim.run(3000)
    ce, we = eng.get_hills()
    co, wo = orc.get_hills()
    assert len(we) == len(wo) == 30
    assert we[0] == 0.05
    assert np.allclose(we, wo, rtol=1e-9)
    assert np.allclose(ce, co, rtol=1e-9)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "fixf", "biase", "totale"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k
    assert np.allclose(eng.get_vr(), orc.get_vr(), rtol=1e-9)
    # FES estimate on the reference's plot grid (bias/gaus.py:365-393, io/plot.py:517-519)
    ga = eng.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40)
    gb = orc.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40)
    assert scaled_err(ga, gb) < 1e-9


@pytest.mark.parametrize("shared", [True, False])
def test_mtd_multi_walker_external_noise(shared):
    """New semantics: M walkers deposit together; shared table (multiple-walker metadynamics) or one
    table per walker (M independent copies of the reference run)."""
    M = 12
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
               true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0,
               temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=50,
               mtd_biasf=10.0, steps=600, mtd_shared=shared)
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=64)
    noise = noise_table(9, 600, 5, M)
    rng = np.random.RandomState(1)
    x0 = spec.x0.reshape(-1, 1) + rng.normal(scale=0.5, size=(5, M))
    for sim in (eng, orc):
        sim.set_state(x0, rng.normal(scale=0.0, size=(5, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(600)
    for wk in (0, M - 1):
        ce, we = eng.get_hills(wk)
        co, wo = orc.get_hills(wk)
        assert len(we) == len(wo) == (12 * M if shared else 12)
        assert np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "biase"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k


@pytest.mark.parametrize("case", ["XmY_shared", "Proj5dto1d_per_walker"])
def test_mtd_adapsig_geo_external_noise(case):
    """adapsig_geo with a 1-D colvar (gaus.py:162-165,189-192,263-288): every hill carries its own variance
    J.J^T (first hill of a table) / J.J^T sigma^2 taken at the deposition point.  Hill tables incl. variances,
    states, VR, the FES grid (gaus.py:428-434) and point evaluations must equal the oracle, which is pinned to
    the reference's own adaptive runs (tests/golden/trajectories.json)."""
    M = 10
    adap = dict(integrate="langevin", gamma=0.1, dt=1.0, temperature=300, mass=1.0, biases=["mtd"], mtd_w=0.05,
                mtd_sigma=2.0, mtd_dep_freq=50, mtd_biasf=10.0, mtd_adapsig=True, mtd_adapsig_geo=True, steps=600)
    if case == "XmY_shared":
        cfg = dict(adap, ndim=2, potential="Mueller2d", colvar_name="XmY", x0=[22.0, 24.0], mtd_shared=True)
    else:
        cfg = dict(adap, ndim=5, potential="Mueller5d", colvar_name="Proj5dto1d", true_colvar_name="Proj5dto2d",
                   x0=[12.0] * 5, mtd_shared=False)
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=64)
    nd = spec.ndim
    noise = noise_table(21, 600, nd, M)
    rng = np.random.RandomState(2)
    x0 = spec.x0.reshape(-1, 1) + rng.normal(scale=0.5, size=(nd, M))
    for sim in (eng, orc):
        sim.set_state(x0, np.zeros((nd, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(600)
    shared = cfg["mtd_shared"]
    for wk in (0, M - 1):
        (ce, we), (co, wo) = eng.get_hills(wk), orc.get_hills(wk)
        se, so = eng.get_hill_sigma2(wk), orc.get_hill_sigma2(wk)
        assert len(we) == len(wo) == len(se) == (12 * M if shared else 12)
        assert np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9) and np.allclose(se, so, rtol=1e-12)
    s0 = eng.get_hill_sigma2(0)
    if case == "XmY_shared":
        assert np.all(s0[:M] == 2.0) and np.all(s0[M:] == 8.0)       # J = (1, -1): 2, then 2 * sigma^2
    else:
        assert np.all(np.abs(s0[:1] - 1.0) < 1e-6) and np.all(np.abs(s0[1:] - 4.0) < 1e-5) and s0[1] != s0[2]
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "fixf", "biase", "totale"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k
    assert np.allclose(eng.get_vr(), orc.get_vr(), rtol=1e-9)
    lo = float(min(ce.min(), co.min())) - 5.0
    assert scaled_err(eng.fes_grid(lo, 0.25, 120), orc.fes_grid(lo, 0.25, 120)) < 1e-9
    pts = x0 + 0.3
    ea, eb = eng.eval_at(pts), orc.eval_at(pts)
    assert scaled_err(ea["Vb"], eb["Vb"]) < 1e-11 and scaled_err(ea["Fb"], eb["Fb"]) < 1e-11
    if shared:  # checkpoint / restore carries the variances
        ck = eng.checkpoint()
        eng2 = Engine(spec.to_struct())
        eng2.restore(ck)
        assert np.array_equal(eng2.get_hill_sigma2(0), s0)
        assert scaled_err(eng2.eval_at(pts)["Vb"], ea["Vb"]) < 1e-14


@pytest.mark.parametrize("case", ["1d_shared_f64", "2d_xmy_per_walker_f64", "1d_shared_f32"])
def test_mtd_adapsig_history_external_noise(case):
    """adapsig without adapsig_geo (gaus.py:166-183, 193-212): the step kernel keeps the Stat.colv window of every
    walker in a device ring (sampled every stat_freq steps, launch boundaries anywhere), the deposition takes the
    exponentially weighted variance of it.  Many walkers, shared and per-walker tables, against the oracle (pinned to
    the reference's own runs by tests/golden/trajectories.json); checkpoint / resume carries the window."""
    M, n = 12, 700
    hist = dict(integrate="langevin", gamma=0.1, dt=1.0, temperature=300, mass=1.0, biases=["mtd"], mtd_w=0.02,
                mtd_sigma=2.0, mtd_dep_freq=30, mtd_biasf=10.0, mtd_adapsig=True, mtd_adapsig_geo=False, steps=n)
    if case.startswith("1d"):
        cfg = dict(hist, ndim=1, potential="DoubleWell1d", x0=30.0, mtd_shared=True, stat_freq=3)
    else:  # a (1, 1)-shaped colvar: the reference cannot run this one (dcol.T.dot fails), the engine and oracle can
        cfg = dict(hist, ndim=2, potential="Mueller2d", colvar_name="XmY", x0=[22.0, 24.0], mtd_shared=False)
    f32 = case.endswith("f32")
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=37,
                               precision="f32" if f32 else "f64")
    nd = spec.ndim
    noise = noise_table(33, n, nd, M)
    rng = np.random.RandomState(4)
    x0 = spec.x0.reshape(-1, 1) + rng.normal(scale=0.5, size=(nd, M))
    for sim in (eng, orc):
        sim.set_state(x0, np.zeros((nd, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(400)
    ck = eng.checkpoint()
    for sim in (eng, orc):
        sim.run(n - 400)
    tol = 5e-3 if f32 else TRAJ_TOL
    ndep = n // 30 + 1
    for wk in (0, M - 1):
        (ce, we), (co, wo) = eng.get_hills(wk), orc.get_hills(wk)
        se, so = eng.get_hill_sigma2(wk), orc.get_hill_sigma2(wk)
        assert len(we) == len(wo) == len(se) == (ndep * M if cfg["mtd_shared"] else ndep)
        assert np.allclose(se, so, rtol=1e-3 if f32 else 1e-8)
        assert np.allclose(we, wo, rtol=tol) and np.allclose(ce, co, rtol=tol)
    s0 = eng.get_hill_sigma2(0)
    assert np.all(s0[:M if cfg["mtd_shared"] else 1] == 1.0)    # one sample: |dcol| < 1e-10 -> variance 1
    assert len(np.unique(s0)) > ndep // 2                       # later hills: each its own variance
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "biase") if f32 else ("x", "v", "fixf", "biase"):  # fp32: narrow hills amplify rounding in fixf
        assert scaled_err(a[k], b[k]) < tol, k
    if not f32:
        ring, ns = eng.get_cv_history()
        assert ns == 1 + n // spec.stat_freq and ring.shape == (100, M)
        # resume from step 400 on a fresh engine: same hills, same state
        eng2 = Engine(spec.to_struct())
        eng2.restore(ck)
        eng2.set_noise(noise)
        eng2.run(n - 400)
        a2 = eng2.get_state()
        for k in ("x", "v", "biase"):
            assert np.array_equal(a2[k], a[k]), k
        assert np.array_equal(eng2.get_hill_sigma2(0), s0)


FES = load_golden("fes_projection.json")


@pytest.mark.parametrize("name", sorted(FES))
def test_fes_grid_matches_reference_projection(name):
    """nds_fes_grid against Gaussian.projection of the reference itself (gaus.py:365-438 on the plot grid of
    io/plot.py:483-488), incremental calls included."""
    rec = FES[name]
    spec = spec_from(rec["config"], precision="f64")
    centers, heights = np.array(rec["hills_centers"]), np.array(rec["hills_heights"])
    args = []
    for (b0, b1), inc in zip(rec["boundary"], rec["increment"]):
        args += [b0, inc, len(np.arange(b0, b1, inc))]
    for n, want in zip(rec["hills_at_call"], rec["maps"]):
        eng = Engine(spec.to_struct())
        eng.set_hills(centers[:n], heights[:n], sigma2=rec["hills_sigma2"][:n] if "hills_sigma2" in rec else None)
        g = eng.fes_grid(*args)
        want = np.array(want).reshape(rec["shape"])
        assert g.reshape(want.shape) == pytest.approx(want, rel=1e-11, abs=1e-300)


def test_committor_external_noise():
    """Basin id and exit step must be identical (same inclusive fp64 comparisons, committor.py:127-140)."""
    g = load_golden("committor.json")
    cfg = dict(g["config"], steps=4000)
    M = 256
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=500)
    noise = noise_table(21, 4000, 2, M)
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    v0 = np.random.RandomState(2).normal(scale=0.0157, size=(2, M))
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        sim.set_noise(noise)
        sim.begin()
        sim.run(4000)
    be, se = eng.get_commit()
    bo, so = orc.get_commit()
    assert np.array_equal(be, bo)
    assert np.array_equal(se, so)
    assert (be >= 0).sum() > 10  # the case does commit
    assert eng.count_alive() == int((be < 0).sum())
    a, b = eng.get_state(), orc.get_state()
    assert scaled_err(a["x"], b["x"]) < TRAJ_TOL


@pytest.mark.parametrize("which", ["shots", "diffusion"])
def test_committor_shots_of_the_reference_with_its_own_noise(which):
    """The reference's 24 committor shots (tests/golden/committor.json, one seed each) run as ONE 24-walker
    ensemble: every walker gets the Gaussian numbers RandomState(seed) handed to its shot.  Basin ids, exit steps
    and final positions must equal the reference's, and uncommitted shots report basin -1 / all steps."""
    g = load_golden("committor.json")
    if which == "diffusion":   # started inside basin 0 with diffusion_time = 37: exit at step 38 (committor.py:100)
        g = g["diffusion"]
    shots = g["shots"]
    cfg = dict(g["config"])
    nsteps, M = cfg["steps"], len(shots)
    spec = spec_from(cfg, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=1000)
    table = np.zeros((nsteps, 2, M))
    v0 = np.zeros((2, M))
    for w, sh in enumerate(shots):
        z0, t = reference_normals(sh["seed"], 2, nsteps, 2)
        table[:, :, w] = t[:, :, 0]
        v0[:, w] = sh["v0"]
        assert np.allclose(z0 * np.sqrt(kB_ * cfg["temperature"] / (104.23315626367025 * cfg["mass"])), sh["v0"], rtol=1e-13)
    eng = Engine(spec.to_struct())
    eng.set_state(np.repeat(spec.x0.reshape(-1, 1), M, axis=1), v0)
    eng.set_noise(table)
    eng.begin()
    eng.run(nsteps)
    basin, exit_step = eng.get_commit()
    assert list(basin) == [sh["basin"] for sh in shots]
    assert list(exit_step) == [sh["step"] for sh in shots]
    if which == "shots":
        assert len(set(basin)) == 3                  # both basins and uncommitted shots occur
    x = eng.get_state()["x"]
    assert scaled_err(x.T, np.array([sh["x"] for sh in shots])) < 1e-8


def test_umbrella_windows_per_walker():
    """Per-walker (k, r0): walker w with window (k_w, r0_w) must equal a single-walker run of that window."""
    base = EXT_CASES["umb_5d"][0]
    M, nsteps = 6, 500
    cfg = dict(base, umb_per_walker=True)
    spec, eng, _ = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL)
    noise = noise_table(4, nsteps, 10, M)
    r0 = np.array([[10 + 4 * w for w in range(M)], [30 - 3 * w for w in range(M)]], dtype=float)
    k = np.array([[0.1 + 0.01 * w for w in range(M)], [0.1] * M])
    x0 = np.stack([r0[0], np.zeros(M), r0[1], np.zeros(M), np.full(M, 100.0)])
    eng.set_state(x0, np.zeros((5, M)))
    eng.set_walker_bias(k, r0)
    eng.set_noise(noise)
    eng.begin()
    eng.run(nsteps)
    a = eng.get_state()
    for w in range(M):
        c1 = dict(base, umb_k=list(k[:, w]), umb_r0=list(r0[:, w]))
        s1 = spec_from(c1, n_walkers=1, noise_mode=_abi.NDS_NOISE_EXTERNAL).to_struct()
        o = oracle.OracleSim(s1, seeds=[0])
        o.set_state(x0[:, w:w + 1], np.zeros((5, 1)))
        o.set_noise(noise[:, :, w:w + 1])
        o.begin()
        o.run(nsteps)
        b = o.get_state()
        assert scaled_err(a["x"][:, w], b["x"][:, 0]) < TRAJ_TOL
        assert scaled_err(a["biase"][w:w + 1], b["biase"]) < TRAJ_TOL


def test_frames_match_oracle():
    """Frames recorded on the device (io/dump.py content) equal the oracle's trajectory rows."""
    cfg = dict(MD2D, integrate="langevin", gamma=0.1, dump=True, dump_freq=10, stat_freq=1, steps=200,
               dump_walkers=3)
    M = 5
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=64)
    noise = noise_table(8, 200, 2, M)
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    v0 = np.random.RandomState(3).normal(scale=0.01, size=(2, M))
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        sim.set_noise(noise)
        sim.begin()
    eng.run(200)
    traj = orc.run(200, traj_every=1)
    fr = orc.split_frame(traj)
    frames = eng.drain_frames()
    names = eng.frame_fields()
    steps = frames[:, names.index("step"), 0].astype(int)
    # dump steps 0,10,...,200 and the lagged stat samples 9,19,...,199
    assert list(steps) == sorted([0] + list(range(10, 201, 10)) + list(range(9, 200, 10)))
    for i, s in enumerate(steps):
        if s == 0:
            assert np.allclose(frames[i, names.index("x0"), :], x0[0, :3])
            continue
        for d in range(2):
            assert scaled_err(frames[i, names.index(f"x{d}"), :], fr["x"][s - 1, d, :3]) < TRAJ_TOL
            assert scaled_err(frames[i, names.index(f"f{d}"), :], fr["f"][s - 1, d, :3]) < TRAJ_TOL
        assert scaled_err(frames[i, names.index("pe"), :], fr["pe"][s - 1, :3]) < TRAJ_TOL
        assert scaled_err(frames[i, names.index("T"), :], fr["T"][s - 1, :3]) < TRAJ_TOL
    assert eng.frames_pending() == 0


def test_launch_boundaries_do_not_change_results():
    """K (steps per launch) is an execution detail: Philox noise is indexed by the global step."""
    cfg = dict(MD2D, integrate="langevin", gamma=0.1, seed=123)
    outs = []
    for K in (1000, 37, 1):
        spec = spec_from(cfg, n_walkers=64, steps_per_launch=K)
        eng = Engine(spec.to_struct())
        eng.set_state(np.repeat(spec.x0.reshape(-1, 1), 64, axis=1))
        eng.begin()
        eng.run(300)
        outs.append(eng.get_state())
    for o in outs[1:]:
        assert scaled_err(o["x"], outs[0]["x"]) < 1e-11
        assert scaled_err(o["v"], outs[0]["v"]) < 1e-11


def test_walker_offset_makes_sharding_invisible():
    """Rank r owning walkers [off, off+M) reproduces the same walkers of a single-rank run."""
    cfg = dict(MD2D, integrate="2nd-langevin", gamma=0.01, seed=5)
    spec = spec_from(cfg, n_walkers=96)
    full = Engine(spec.to_struct())
    full.set_state(np.repeat(spec.x0.reshape(-1, 1), 96, axis=1))
    full.begin()
    full.run(200)
    a = full.get_state()
    spec2 = spec_from(cfg, n_walkers=32, walker_offset=64, n_walkers_global=96)
    part = Engine(spec2.to_struct())
    part.set_state(np.repeat(spec.x0.reshape(-1, 1), 32, axis=1))
    part.begin()
    part.run(200)
    b = part.get_state()
    assert np.array_equal(a["x"][:, 64:], b["x"])
    assert np.array_equal(a["v"][:, 64:], b["v"])


def test_fp32_one_step_close_to_fp64():
    """fp32 kernels (fast MUFU forms) stay within single-precision distance of the fp64 path per step."""
    rng = np.random.RandomState(0)
    M = 4096
    x0 = np.stack([rng.uniform(10, 40, M), rng.uniform(5, 35, M)])
    v0 = rng.normal(scale=0.0157, size=(2, M))
    outs = {}
    for prec in ("f64", "f32"):
        spec = spec_from(dict(MD2D, integrate="nve"), n_walkers=M, precision=prec)
        eng = Engine(spec.to_struct())
        eng.set_state(x0, v0)
        eng.begin()
        eng.run(1)
        outs[prec] = eng.get_state()
    assert scaled_err(outs["f32"]["x"], outs["f64"]["x"]) < 1e-6
    assert np.max(np.abs(outs["f32"]["f"] - outs["f64"]["f"])) < 1e-5 * np.max(np.abs(outs["f64"]["f"]))
    assert np.max(np.abs(outs["f32"]["pe"] - outs["f64"]["pe"])) < 2e-6


@pytest.mark.parametrize("integrate", ["langevin", "2nd-langevin"])
def test_mtd_block_many_tiles(integrate):
    """The TMA-tiled shared-hill kernel (mtd_block): table spanning several 1024-record tiles, walkers
    not a multiple of the per-CTA count, against the oracle with external noise."""
    M, nsteps = 45, 500
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
               true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate=integrate, gamma=0.1, dt=1.0,
               temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=10,
               mtd_biasf=10.0, steps=nsteps, mtd_shared=True)
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=100)
    assert eng.kernel_name.startswith("mtd_block")
    nn = 5 if integrate == "langevin" else 10
    noise = noise_table(13, nsteps, nn, M)
    rng = np.random.RandomState(2)
    x0 = spec.x0.reshape(-1, 1) + rng.normal(scale=0.5, size=(5, M))
    v0 = rng.normal(scale=0.01, size=(5, M))
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        sim.set_noise(noise)
        sim.begin()
        sim.run(nsteps)
    ce, we = eng.get_hills()
    co, wo = orc.get_hills()
    assert len(we) == len(wo) == 50 * M > 2048
    assert np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "f", "fixf", "biase", "pe", "totale"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k
    assert np.allclose(eng.get_vr(), orc.get_vr(), rtol=1e-9)


@pytest.mark.parametrize("case", ["XmY_2d", "Proj5dto1d_umb", "DoubleWell1d"])
def test_mtd_block_one_dimensional_colvar(case):
    """Shared hill table over a 1-D colvar: the records become (c, 0, w, log2 w) and the TMA-tiled 2-D hill pass
    runs with c1 = 0 and 1/sigma_1^2 = 0.  Against the oracle with external noise (fp64), plus fp32 vs fp64 of the
    bias on the final table, FES grid and hill round trip."""
    M, nsteps = 37, 400
    base = dict(mass=1.0, integrate="langevin", gamma=0.1, dt=1.0, temperature=300, mtd_w=0.05, mtd_sigma=1.5,
                mtd_dep_freq=10, mtd_biasf=10.0, steps=nsteps, mtd_shared=True)
    if case == "XmY_2d":
        cfg = dict(base, ndim=2, potential="Mueller2d", colvar_name="XmY", x0=[22.0, 24.0], biases=["mtd"])
    elif case == "DoubleWell1d":   # 1-D system: multiple-walker metadynamics across the double-well barrier
        cfg = dict(base, ndim=1, potential="DoubleWell1d", x0=30.0, biases=["mtd"], temperature=1000)
    else:
        cfg = dict(base, ndim=5, potential="Mueller5d", colvar_name="Proj5dto1d", true_colvar_name="Proj5dto2d",
                   x0=[12.0] * 5, biases=["umb", "mtd"], umb_k=0.02, umb_r0=[15.0])
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=100)
    assert eng.kernel_name.startswith("mtd_block"), eng.kernel_name
    nd = spec.ndim
    noise = noise_table(17, nsteps, nd, M)
    rng = np.random.RandomState(4)
    x0 = spec.x0.reshape(-1, 1) + rng.normal(scale=0.5, size=(nd, M))
    for sim in (eng, orc):
        sim.set_state(x0, np.zeros((nd, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(nsteps)
    (ce, we), (co, wo) = eng.get_hills(), orc.get_hills()
    assert ce.shape == co.shape == (40 * M, 1) and 40 * M > 1024
    assert np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "f", "fixf", "colv", "biase", "pe", "totale"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k
    assert np.allclose(eng.get_vr(), orc.get_vr(), rtol=1e-9)
    lo = float(co.min()) - 4.0
    assert scaled_err(eng.fes_grid(lo, 0.2, 100), orc.fes_grid(lo, 0.2, 100)) < 1e-9
    # fp32 instantiation (MUFU.EX2, packed FP32) on the same table
    spec32 = spec_from(cfg, n_walkers=M, precision="f32")
    e32 = Engine(spec32.to_struct())
    assert e32.kernel_name.startswith("mtd_block")
    e32.set_state(a["x"], a["v"])
    e32.begin()
    e32.set_hills(ce, we)
    e32.set_clock(nsteps + 1, nsteps + 1.0, 41)   # no deposition due: the launch runs on the frozen table
    e32.run(1)
    st = e32.get_state()
    ref = eng.eval_at(st["x"])                   # fp64, generic hill sum, at the positions fp32 ended on
    assert e32.n_hills() == 40 * M
    assert scaled_err(st["biase"], ref["Vb"]) < 2e-5
    assert scaled_err(st["fixf"], ref["Fb"]) < 2e-4


def test_mtd_block_fp32_close_to_fp64():
    """fp32 mtd_block (prescaled exponent, MUFU.EX2) vs the fp64 instantiation on a frozen table."""
    M = 300
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
               true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0,
               temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=1e9,
               mtd_biasf=10.0, steps=10, mtd_shared=True, mtd_capacity=5000)
    rng = np.random.RandomState(4)
    centers = rng.uniform(8, 18, size=(3000, 2))
    heights = rng.uniform(0.01, 0.05, size=3000)
    x0 = 12.0 + rng.normal(scale=1.0, size=(5, M))
    outs = {}
    for prec in ("f64", "f32"):
        spec = spec_from(cfg, n_walkers=M, precision=prec)
        eng = Engine(spec.to_struct())
        assert eng.kernel_name.startswith("mtd_block")
        eng.set_state(x0, np.zeros((5, M)))
        eng.set_hills(centers, heights)
        eng.begin()
        outs[prec] = eng.get_state()
    # begin deposits nothing; the table has the 3000 given hills
    assert np.max(np.abs(outs["f32"]["biase"] - outs["f64"]["biase"])) < 2e-5 * np.max(outs["f64"]["biase"])
    assert np.max(np.abs(outs["f32"]["fixf"] - outs["f64"]["fixf"])) < 1e-4 * np.max(np.abs(outs["f64"]["fixf"]))


def test_two_rank_shared_hills_emulated_on_one_gpu():
    """N > 1 path of shared-hill metadynamics: two engines ("ranks", walkers [0,M) and [M,2M)) driven by
    two host threads on one GPU, their exchange callbacks completing each new hill segment from the
    peer's table (the role of the NCCL allgather), must reproduce one engine with 2M walkers exactly."""
    import threading
    import torch
    from ndsimulator_b200.distributed import _DevSegment

    M, nsteps = 20, 300
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
               true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0,
               temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=25,
               mtd_biasf=10.0, steps=nsteps, mtd_shared=True, seed=77)
    rng = np.random.RandomState(6)
    x0 = 12.0 + rng.normal(scale=0.5, size=(5, 2 * M))
    one = Engine(spec_from(cfg, n_walkers=2 * M).to_struct())
    one.set_state(x0)
    one.begin()
    one.run(nsteps)
    ref = one.get_state()
    ref_hills = one.get_hills()

    engines = [Engine(spec_from(cfg, n_walkers=M, walker_offset=r * M, n_walkers_global=2 * M).to_struct())
               for r in range(2)]
    barrier = threading.Barrier(2)
    seg_ptr = [None, None]

    def make_cb(r):
        def cb(ptr, bytes_per_rank, rank_offset_bytes):
            seg_ptr[r] = ptr
            barrier.wait()
            mine = torch.as_tensor(_DevSegment(seg_ptr[r], 2 * bytes_per_rank), device="cuda")
            peer = torch.as_tensor(_DevSegment(seg_ptr[1 - r], 2 * bytes_per_rank), device="cuda")
            po = (1 - r) * bytes_per_rank
            mine[po:po + bytes_per_rank].copy_(peer[po:po + bytes_per_rank])
            torch.cuda.synchronize()
            barrier.wait()
            return 0
        return cb

    errors = []

    def drive(r):
        try:
            e = engines[r]
            e.set_exchange(make_cb(r))
            e.set_state(x0[:, r * M:(r + 1) * M])
            e.begin()
            e.run(nsteps)
        except Exception as ex:  # pragma: no cover
            errors.append(ex)
            barrier.abort()

    threads = [threading.Thread(target=drive, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    for r in range(2):
        st = engines[r].get_state()
        for k in ("x", "v", "biase"):
            assert np.array_equal(st[k][..., :], ref[k][..., r * M:(r + 1) * M]), (r, k)
        c, h = engines[r].get_hills()
        assert np.array_equal(c, ref_hills[0]) and np.array_equal(h, ref_hills[1])


def test_packed_fp32_kernel_matches_scalar_fp32(monkeypatch):
    """The packed-FP32 (two walkers per thread, FFMA2) C2 kernel draws the same Philox numbers and does
    the same arithmetic as the scalar fp32 kernel: a few steps agree to fp32 rounding."""
    M = 4096
    cfg = dict(MD2D, integrate="langevin", gamma=0.1, seed=99)
    rng = np.random.RandomState(1)
    x0 = np.stack([rng.uniform(15, 30, M), rng.uniform(10, 35, M)])
    outs = {}
    for mode in ("packed", "scalar"):
        if mode == "scalar":
            monkeypatch.setenv("NDS_NO_PACKED", "1")
        spec = spec_from(cfg, n_walkers=M, precision="f32")
        eng = Engine(spec.to_struct())
        assert ("f32x2" in eng.kernel_name) == (mode == "packed")
        eng.set_state(x0)
        eng.begin()
        eng.run(5)
        outs[mode] = eng.get_state()
    assert scaled_err(outs["packed"]["x"], outs["scalar"]["x"]) < 1e-6
    assert scaled_err(outs["packed"]["v"], outs["scalar"]["v"]) < 1e-4
    assert scaled_err(outs["packed"]["pe"], outs["scalar"]["pe"]) < 1e-4


def test_peer_memory_deposit_emulated_on_one_gpu():
    """Fused deposition + exchange over peer pointers: two engines in one process write their new hill
    records directly into both tables (nds_set_peer_tables); the callback is only a barrier.  Must equal
    one engine with 2M walkers bit for bit."""
    import threading

    M, nsteps = 24, 200
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
               true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0,
               temperature=300, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=20,
               mtd_biasf=10.0, steps=nsteps, mtd_shared=True, seed=5, precision="f32")
    rng = np.random.RandomState(8)
    x0 = 12.0 + rng.normal(scale=0.5, size=(5, 2 * M))
    one = Engine(spec_from(cfg, n_walkers=2 * M).to_struct())
    one.set_state(x0)
    one.begin()
    one.run(nsteps)
    ref, ref_hills = one.get_state(), one.get_hills()

    engines = [Engine(spec_from(cfg, n_walkers=M, walker_offset=r * M, n_walkers_global=2 * M).to_struct())
               for r in range(2)]
    ptrs = [e.device_pointers()["hills"] for e in engines]
    barrier = threading.Barrier(2)
    calls = [0, 0]

    def make_cb(r):
        def cb(ptr, nbytes, off):
            assert nbytes == 0 and ptr is None
            calls[r] += 1
            barrier.wait()
            return 0
        return cb

    errors = []

    def drive(r):
        try:
            e = engines[r]
            e.set_peer_tables(ptrs, r)
            e.set_exchange(make_cb(r))
            e.set_state(x0[:, r * M:(r + 1) * M])
            e.begin()
            e.run(nsteps)
        except Exception as ex:  # pragma: no cover
            errors.append(ex)
            barrier.abort()

    threads = [threading.Thread(target=drive, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    assert calls == [10, 10]
    for r in range(2):
        st = engines[r].get_state()
        for k in ("x", "v", "biase"):
            assert np.array_equal(st[k], ref[k][..., r * M:(r + 1) * M]), (r, k)
        c, h = engines[r].get_hills()
        assert np.array_equal(c, ref_hills[0]) and np.array_equal(h, ref_hills[1])


def test_committor_compaction_is_invisible(monkeypatch):
    """Lane compaction between launches (live walkers packed to the front) must not change any result:
    same basins, exit steps and final states as with compaction disabled."""
    g = load_golden("committor.json")
    cfg = dict(g["config"], steps=6000, seed=4)
    M = 20000
    outs = []
    for mode in ("compact", "plain"):
        if mode == "plain":
            monkeypatch.setenv("NDS_NO_COMPACTION", "1")
        spec = spec_from(cfg, n_walkers=M, precision="f32", steps_per_launch=500)
        eng = Engine(spec.to_struct())
        eng.set_state(np.repeat(spec.x0.reshape(-1, 1), M, axis=1))
        eng.begin()
        eng.run(6000)
        basin, exit_step = eng.get_commit()
        outs.append((basin, exit_step, eng.get_state(), eng.count_alive()))
    (b0, e0, s0, a0), (b1, e1, s1, a1) = outs
    assert (b0 >= 0).sum() > 1000
    assert np.array_equal(b0, b1) and np.array_equal(e0, e1) and a0 == a1
    for k in ("x", "v", "colv", "pe", "T"):
        assert np.array_equal(s0[k], s1[k]), k


def test_packed_committor_matches_scalar(monkeypatch):
    """Two shots per thread (packed fp32 committor kernel) against the scalar fp32 committor kernel: same
    Philox numbers, same arithmetic up to FMA contraction, so over a few steps from start points around the
    basin boxes the basins / exit steps agree (a stopped lane must keep the state of its exit step while
    its partner lane keeps integrating)."""
    g = load_golden("committor.json")
    cfg = dict(g["config"], steps=40, seed=11, diffusion_time=3)
    M = 8192
    rng = np.random.RandomState(5)
    x0 = np.stack([rng.uniform(18, 52, M), rng.uniform(0, 37, M)])
    outs = {}
    for mode in ("packed", "scalar"):
        if mode == "scalar":
            monkeypatch.setenv("NDS_NO_PACKED", "1")
        spec = spec_from(cfg, n_walkers=M, precision="f32", steps_per_launch=16)
        eng = Engine(spec.to_struct())
        assert ("f32x2" in eng.kernel_name) == (mode == "packed"), eng.kernel_name
        eng.set_state(x0)
        eng.begin()
        eng.run(40)
        outs[mode] = (eng.get_commit(), eng.get_state(), eng.count_alive())
    (bp, ep), sp, ap = outs["packed"]
    (bs, es), ss, as_ = outs["scalar"]
    assert 500 < (bs >= 0).sum() < M - 500          # both committed and live lanes are exercised
    assert (bp != bs).mean() < 1e-3 and (ep != es).mean() < 1e-3 and abs(ap - as_) <= 8
    same = (bp == bs) & (ep == es)
    for k in ("x", "v", "colv", "pe"):
        assert scaled_err(sp[k][..., same], ss[k][..., same]) < 1e-3, k
    # a committed walker sits inside its basin box, at the state of its exit step
    crit = np.array(g["config"]["criteria"])
    for b in range(crit.shape[0]):
        sel = bp == b
        assert sel.any()
        for k in range(2):
            assert (sp["colv"][k][sel] >= crit[b, k, 0]).all() and (sp["colv"][k][sel] <= crit[b, k, 1]).all()


def test_committor_temperature_zero_minimize():
    """T = 0 committor on the GPU (steepest-descent inner engine, pe < emin criterion): basin ids and exit
    steps identical to the reference's runs, positions / energies within the trajectory tolerance; plus a
    basin-of-attraction map of 2^16 start points against the oracle."""
    g = load_golden("committor_T0.json")
    runs = g["runs"]
    spec = spec_from(g["config"], n_walkers=len(runs), steps_per_launch=500)
    eng = Engine(spec.to_struct())
    assert "minimize" in eng.kernel_name
    eng.set_state(np.array([r["x0"] for r in runs]).T)
    eng.begin()
    eng.run(g["config"]["steps"])
    basin, exit_step = eng.get_commit()
    assert list(basin) == [r["basin"] for r in runs]
    assert list(exit_step) == [r["step"] for r in runs]
    st = eng.get_state()
    assert scaled_err(st["x"].T, np.array([r["x"] for r in runs])) < TRAJ_TOL
    assert scaled_err(st["pe"], np.array([r["pe"] for r in runs])) < TRAJ_TOL
    assert scaled_err(st["colv"].T, np.array([r["colv"] for r in runs])) < TRAJ_TOL
    # map: many start points, GPU vs oracle
    M = 4096
    rng = np.random.RandomState(0)
    x0 = np.stack([rng.uniform(12, 46, M), rng.uniform(2, 36, M)])
    spec = spec_from(dict(g["config"], steps=1500), n_walkers=M)
    eng, orc = Engine(spec.to_struct()), oracle.OracleSim(spec.to_struct(), seeds=np.zeros(M))
    orc.set_threads(oracle.max_threads())
    for sim in (eng, orc):
        sim.set_state(x0)
        sim.begin()
        sim.run(1500)
    bg, sg = eng.get_commit()
    bo, so = orc.get_commit()
    assert (bg >= 0).sum() > 100
    # ulp-level differences can flip a dE >= 0 stop by a step for walkers sitting at a minimum
    assert (bg != bo).mean() < 0.002 and (np.abs(sg - so) > 1).mean() < 0.01


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_f32_host_buffers_roundtrip(precision):
    """nds_set_state_f32 / nds_get_state_f32 move the same state as the fp64 calls (to fp32 rounding)."""
    M = 1000
    spec = spec_from(dict(MD2D, integrate="nve"), n_walkers=M, precision=precision)
    eng = Engine(spec.to_struct())
    rng = np.random.RandomState(0)
    x0 = np.stack([rng.uniform(15, 30, M), rng.uniform(10, 35, M)]).astype(np.float32)
    v0 = rng.normal(scale=0.01, size=(2, M)).astype(np.float32)
    eng.set_state_f32(x0, v0)
    eng.begin()
    a32 = eng.get_state_f32(fields=("x", "v", "colv", "thermo"))
    a64 = eng.get_state()
    assert np.array_equal(a32["x"], x0) and np.array_equal(a32["v"], v0)
    assert np.allclose(a32["pe"], a64["pe"], rtol=1e-6) and np.allclose(a32["colv"], a64["colv"], rtol=1e-6)
    eng.run(10)
    b32, b64 = eng.get_state_f32(), eng.get_state()
    assert np.allclose(b32["x"], b64["x"], rtol=1e-6) and np.allclose(b32["T"], b64["T"], rtol=1e-5)


@pytest.mark.parametrize("case", ["langevin", "mtd", "mtd_per_walker", "rescale"])
def test_checkpoint_resume_is_bit_exact(case):
    """(x, v, clock, hills, VR) restored on a fresh engine continues the run bit for bit: Philox noise is a
    pure function of (seed, global walker id, global step)."""
    M = 64
    if case == "langevin":
        cfg = dict(MD2D, integrate="langevin", gamma=0.1, seed=9)
    elif case == "rescale":
        cfg = dict(MD2D, integrate="rescale", rescale_freq=7.0, seed=9)
    else:
        cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
                   x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0, temperature=300, biases=["mtd"],
                   mtd_w=0.01, mtd_sigma=[2.0, 2.0], mtd_dep_freq=30, mtd_biasf=10.0, steps=400, seed=9,
                   mtd_shared=(case == "mtd"))
    x0 = np.repeat(np.array(cfg["x0"], dtype=float).reshape(-1, 1), M, axis=1) + \
        np.random.RandomState(0).normal(scale=0.5, size=(len(cfg["x0"]), M))
    spec = spec_from(cfg, n_walkers=M, steps_per_launch=50)
    a = Engine(spec.to_struct())
    a.set_state(x0)
    a.begin()
    a.run(400)
    ref = a.get_state()
    b = Engine(spec.to_struct())
    b.set_state(x0)
    b.begin()
    b.run(170)
    ck = b.checkpoint()
    c = Engine(spec.to_struct())
    c.restore(ck)
    c.run(230)
    got = c.get_state()
    assert c.step == 400 and c.time == a.time
    for k in ("x", "v", "pe", "biase"):
        assert np.array_equal(got[k], ref[k]), k
    if case.startswith("mtd"):
        for wk in (0, M - 1):
            assert np.array_equal(c.get_hills(wk)[1], a.get_hills(wk)[1])
            assert np.array_equal(c.get_hills(wk)[0], a.get_hills(wk)[0])
    assert c.count_nonfinite() == 0


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_committor_checkpoint_resume(precision):
    """A committor ensemble (alive flags, basins, exit steps, frozen states of stopped walkers; lane compaction
    on both sides) restored on a fresh engine finishes exactly like the uninterrupted run."""
    g = load_golden("committor.json")
    cfg = dict(g["config"], steps=6000, seed=21)
    M = 4096
    spec = spec_from(cfg, n_walkers=M, precision=precision, steps_per_launch=500)
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    a = Engine(spec.to_struct())
    a.set_state(x0); a.begin(); a.run(6000)
    b = Engine(spec.to_struct())
    b.set_state(x0); b.begin(); b.run(2500)
    ck = b.checkpoint()
    assert 0 < (ck["alive"] == 0).sum() < M
    c = Engine(spec.to_struct())
    c.restore(ck)
    assert c.count_alive() == int(ck["alive"].sum())
    c.run(3500)
    (ba, ea), (bc, ec) = a.get_commit(), c.get_commit()
    assert np.array_equal(ba, bc) and np.array_equal(ea, ec) and a.count_alive() == c.count_alive()
    sa, sc = a.get_state(), c.get_state()
    for k in ("x", "v", "colv", "pe"):
        assert np.array_equal(sa[k], sc[k]), k


def test_walker_results_do_not_depend_on_ensemble_size():
    """Tail handling of every launch shape: walker w follows the same Philox stream whatever M is (odd M runs the
    scalar kernel, even M the packed one; partial CTAs, a single thread, M = 1)."""
    cfg = dict(MD2D, integrate="langevin", gamma=0.1, seed=31)
    ref = None
    for M in (1024, 1, 2, 3, 127, 128, 129, 130, 257, 1022):
        spec = spec_from(cfg, n_walkers=M, precision="f32", n_walkers_global=1024)
        eng = Engine(spec.to_struct())
        rng = np.random.RandomState(0)
        x0 = np.stack([rng.uniform(18, 30, 1024), rng.uniform(12, 32, 1024)])[:, :M]
        eng.set_state(x0)
        eng.begin()
        eng.run(20)
        st = eng.get_state()
        assert eng.count_nonfinite() == 0
        if ref is None:
            ref = st
            continue
        for k in ("x", "v", "pe", "T"):
            assert scaled_err(st[k], ref[k][..., :M]) < 2e-5, (M, k)


def test_second_begin_keeps_hills_and_deposition_count():
    """NDRun.begin() restarts step and time only (ndrun.py:199-200): Gaussian._history and _dep_count survive, so after
    a second begin() nothing is deposited until the restarted clock passes the old count.  Engine and oracle agree."""
    M = 4
    cfg = dict(MD2D, integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=50,
               mtd_biasf=10.0, steps=600, mtd_shared=False)
    spec, eng, orc = make_pair(cfg, M=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=64)
    noise = noise_table(5, 300, 2, M)
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    for sim in (eng, orc):
        sim.set_state(x0, np.zeros((2, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(170)            # hills at t = 0, 50, 100, 150
        sim.begin()
        sim.run(300)            # dep_count = 4: the next hill comes at t = 200 of the restarted clock, then 250
    (ce, we), (co, wo) = eng.get_hills(0), orc.get_hills(0)
    assert len(we) == len(wo) == 6
    assert np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "biase"):
        assert scaled_err(a[k], b[k]) < TRAJ_TOL, k
