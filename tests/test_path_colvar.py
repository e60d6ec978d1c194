"""Path / Path5dto2d colvars (reference colvars/path.py; "next" row of SURVEY.md section 8f): the
reference's own known-answer values (tests/unit/test_colvar_path.py:18,43,65), fixtures generated from
the reference, finite-difference Jacobians (the reference's own test idiom), and a trajectory."""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import load_golden, spec_from, scaled_err

G = load_golden("path_colvar.json")
CASES = ["kat_1d", "kat_2d", "kat_file", "path5dto2d_file"]
POT = {1: "DoubleWell1d", 2: "Mueller2d", 5: "Mueller5d"}


def case_cfg(rec, name):
    nd = rec["ndim"]
    cfg = dict(ndim=nd, mass=1.0, potential=POT[nd], x0=rec["points"][0], integrate="nve",
               colvar_name="Path5dto2d" if name.startswith("path5") else "Path",
               colvar_ref=rec["ref"], colvar_ref_ind=rec["ref_ind"], colvar_sigma=rec["sigma"])
    if nd == 5:
        cfg["true_colvar_name"] = "Proj5dto2d"
    return cfg


@pytest.mark.parametrize("name", CASES)
def test_oracle_path_matches_reference(name):
    rec = G[name]
    spec = spec_from(case_cfg(rec, name))
    o = oracle.OracleSim(spec.to_struct(), seeds=[0])
    o.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
    r = o.eval_at(np.array(rec["points"]).T)
    assert np.allclose(r["colv"][0], rec["value"], rtol=1e-12)
    assert np.allclose(r["J"][0].T, np.array(rec["jac"]), rtol=1e-10, atol=1e-16)
    if "expect" in rec:  # the reference's unit-test tolerance
        assert abs(r["colv"][0][0] - rec["expect"]) < 1e-6
    # analytic Jacobian vs central finite difference (test_colvar_path.py:24-30)
    x = np.array(rec["points"][1])
    eps = 1e-3
    for d in range(rec["ndim"]):
        dx = np.zeros(rec["ndim"])
        dx[d] = eps
        pp = o.eval_at((x + dx).reshape(-1, 1))["colv"][0][0]
        pm = o.eval_at((x - dx).reshape(-1, 1))["colv"][0][0]
        ana = o.eval_at(x.reshape(-1, 1))["J"][0][d][0]
        assert abs((pp - pm) / 2 / eps - ana) < 1e-6


def test_oracle_umbrella_on_path_trajectory():
    rec = G["umb_on_path_nve"]
    spec = spec_from(rec["config"])
    o = oracle.OracleSim(spec.to_struct(), seeds=[rec["config"]["seed"]])
    o.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
    o.set_state(spec.x0.reshape(-1, 1))
    o.begin()
    assert np.allclose(o.get_state()["v"][:, 0], rec["v0"], rtol=1e-15)
    traj = o.run(rec["steps"][-1], traj_every=rec["sample_every"])
    fr = o.split_frame(traj)
    for key in ("x", "v", "fixf", "colv"):
        assert scaled_err(fr[key][:, :, 0], np.array(rec[key]).reshape(len(rec["steps"]), -1)) < 1e-9, key
    assert scaled_err(fr["biase"][:, 0], np.array(rec["biase"])) < 1e-9


def test_path_config_errors():
    from ndsimulator_b200 import config
    with pytest.raises(TypeError):
        config.parse(dict(ndim=2, potential="Mueller2d", x0=[1.0, 2.0], colvar_name="Path"))
    with pytest.raises(ValueError):
        config.parse(dict(ndim=2, potential="Mueller2d", x0=[1.0, 2.0], colvar_name="Path",
                          colvar_ref=[[1.0], [2.0]], colvar_ref_ind=[0, 1]))
    with pytest.raises(NotImplementedError):
        config.parse(dict(ndim=2, potential="Mueller2d", x0=[1.0, 2.0], true_colvar_name="Path"))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_path_matches_reference_and_oracle(name):
    from ndsimulator_b200.engine import Engine
    rec = G[name]
    spec = spec_from(case_cfg(rec, name))
    st = spec.to_struct()
    eng = Engine(st)
    eng.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
    r = eng.eval_at(np.array(rec["points"]).T)
    assert np.allclose(r["colv"][0], rec["value"], rtol=1e-11)
    assert np.allclose(r["J"][0].T, np.array(rec["jac"]), rtol=1e-9, atol=1e-15)
    if "expect" in rec:
        assert abs(r["colv"][0][0] - rec["expect"]) < 1e-6
    orc = oracle.OracleSim(st, seeds=[0])
    orc.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
    rng = np.random.RandomState(3)
    base = np.array(rec["points"][0]).reshape(-1, 1)
    more = base + rng.normal(scale=2.0, size=(rec["ndim"], 500))
    a, b = eng.eval_at(more), orc.eval_at(more)
    assert scaled_err(a["colv"], b["colv"]) < 1e-11
    assert np.max(np.abs(a["J"] - b["J"])) < 1e-9 * max(np.max(np.abs(b["J"])), 1e-30) + 1e-15


@pytest.mark.gpu
def test_gpu_umbrella_on_path_trajectory():
    """NDRun-level: umbrella restraint on a path variable, fp64, against the reference's trajectory."""
    from ndsimulator_b200.engine import Engine
    rec = G["umb_on_path_nve"]
    spec = spec_from(rec["config"], steps_per_launch=rec["sample_every"])
    eng = Engine(spec.to_struct())
    eng.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
    eng.set_state(spec.x0.reshape(-1, 1), np.array(rec["v0"]).reshape(-1, 1))
    eng.begin()
    got = {k: [] for k in ("x", "v", "fixf", "colv", "biase")}
    prev = 0
    for s in rec["steps"]:
        eng.run(s - prev)
        prev = s
        stt = eng.get_state()
        for k in got:
            got[k].append(np.atleast_1d(stt[k][..., 0]).copy())
    for k in got:
        assert scaled_err(np.array(got[k]).reshape(len(rec["steps"]), -1),
                          np.array(rec[k]).reshape(len(rec["steps"]), -1)) < 1e-9, k


@pytest.mark.gpu
def test_gpu_metadynamics_on_path5dto2d_external_noise():
    """1-D metadynamics along Path5dto2d (colvardim 1 -> generic hill loop), M walkers vs the oracle."""
    from ndsimulator_b200 import _abi
    from ndsimulator_b200.engine import Engine
    rec = G["path5dto2d_file"]
    M, nsteps = 6, 400
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", true_colvar_name="Proj5dto2d", colvar_name="Path5dto2d",
               colvar_ref=rec["ref"], colvar_ref_ind=rec["ref_ind"], colvar_sigma=2.0, x0=[20.0, 0.0, 27.0, 0.0, 10.0],
               integrate="langevin", gamma=0.1, dt=1.0, temperature=300, biases=["mtd"], mtd_w=0.02,
               mtd_sigma=[0.05], mtd_dep_freq=40, mtd_biasf=8.0, steps=nsteps)
    spec = spec_from(cfg, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=100)
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    noise = np.random.RandomState(1).normal(size=(nsteps, 5, M))
    x0 = spec.x0.reshape(-1, 1) + np.random.RandomState(2).normal(scale=0.5, size=(5, M))
    for sim in (eng, orc):
        sim.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
        sim.set_state(x0, np.zeros((5, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(nsteps)
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "v", "colv", "biase", "fixf"):
        assert scaled_err(a[k], b[k]) < 1e-9, k
    ce, we = eng.get_hills()
    co, wo = orc.get_hills()
    assert len(we) == 10 * M and np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9)
