"""Differential fuzzing of the CUDA path against the oracle: seeded random VALID configurations (potential x colvar x
true colvar x integrator x bias list x frames / launch split), fp64, noise supplied by the caller, a few hundred
steps -- every state array, the hill table and the frames must agree within the trajectory tolerance.  The sample is
fixed by the seed, so a failure names a reproducible configuration."""
import numpy as np
import pytest

from ndsimulator_b200 import _abi
from ndsimulator_b200.engine import Engine
from oracle import oracle
from tests.helpers import scaled_err, spec_from

pytestmark = pytest.mark.gpu

POTS = {1: ["DoubleWell1d", "Flat1d", "HarmonicQuartic"],
        2: ["Mueller2d", "DoubleWell", "ThreeHole2d", "Gaussian2d", "Gaussian", "Flat2d", "HarmonicQuartic"],
        5: ["Mueller5d", "Mueller5dshl", "ThreeHole5d", "HarmonicQuartic"]}
CVS = {1: ["Original"], 2: ["Original", "ExtractX", "ExtractY", "XmY", "XpY", "Rotate2d", "Proj2dto1d"],
       5: ["Original", "Proj5dto1dx0", "Proj5dto1dx2", "Proj5dto1d", "Proj5dto2d", "Proj5dto2dv2"]}
TCVS2 = {2: ["Original", "Rotate2d"], 5: ["Proj5dto2d", "Proj5dto2dv2"]}  # 2-D true colvars
X0 = {1: [28.0], 2: [22.0, 24.0], 5: [14.0, 2.0, 20.0, 1.0, 12.0]}


def random_case(rng):
    nd = int(rng.choice([1, 2, 5]))
    pot = str(rng.choice(POTS[nd]))
    cfg = dict(ndim=nd, mass=float(rng.choice([1.0, 2.5])), potential=pot, temperature=float(rng.choice([150, 300, 600])),
               dt=float(rng.choice([0.5, 1.0])), method="md", steps=400)
    cfg["x0"] = X0[nd][0] if nd == 1 else list(X0[nd])
    if pot == "HarmonicQuartic":
        cfg.update(potential_k=[0.02] * nd, potential_q=[1e-4] * nd, potential_center=[20.0] * nd)
    if pot in ("ThreeHole2d", "ThreeHole5d"):
        cfg["x0"] = [7.0, 9.0] if nd == 2 else [7.0, 1.0, 9.0, 1.0, 3.0]
        cfg["temperature"] = 3000.0
    cfg["colvar_name"] = str(rng.choice(CVS[nd]))
    if nd in TCVS2:
        cfg["true_colvar_name"] = str(rng.choice(TCVS2[nd]))
    integ = str(rng.choice(["langevin", "2nd-langevin", "nve", "rescale"]))
    cfg["integrate"] = integ
    cfg["gamma"] = float(rng.choice([0.0, 0.01, 0.1, 1.0]))
    if integ == "rescale":
        cfg["rescale_freq"] = float(rng.choice([3, 7.5]))
    biases = [b for b in ("umb", "mtd", "pe") if rng.rand() < 0.4]
    cd = {"Original": nd, "Rotate2d": 2, "Proj5dto2d": 2, "Proj5dto2dv2": 2}.get(cfg["colvar_name"], 1)
    if "umb" in biases:
        n = int(rng.choice([1, 2]))
        cfg.update(umb_n=n)
        for i in range(n):
            cfg[f"umb_{i}_k"] = float(rng.choice([0.01, 0.1]))
            cfg[f"umb_{i}_r0"] = [float(v) for v in rng.uniform(10, 30, cd)]
    if "mtd" in biases:
        cfg.update(mtd_w=float(rng.choice([0.01, 0.05])), mtd_sigma=[float(rng.choice([1.0, 2.0]))] * cd,
                   mtd_dep_freq=float(rng.choice([10, 25, 37.5])), mtd_biasf=float(rng.choice([5.0, 10.0])),
                   mtd_shared=bool(rng.rand() < 0.5))
        if cd == 1 and rng.rand() < 0.5:
            cfg.update(mtd_sigma=float(cfg["mtd_sigma"][0]), mtd_adapsig=True, mtd_adapsig_geo=bool(rng.rand() < 0.5))
            if not cfg["mtd_adapsig_geo"]:
                cfg["mtd_dep_freq"] = float(int(cfg["mtd_dep_freq"]))  # tao_d = min(50 dep_freq, 100) must be an integer
                cfg["stat_freq"] = int(rng.choice([1, 3]))
    if "pe" in biases:
        cfg["pe_thred"] = float(rng.choice([-0.5, -0.3, 0.0]))
    cfg["biases"] = biases
    extra = dict(steps_per_launch=int(rng.choice([1, 7, 64, 400])))
    if rng.rand() < 0.3:
        extra.update(dump=True, dump_freq=int(rng.choice([1, 10])), stat_freq=cfg.get("stat_freq", int(rng.choice([1, 5]))),
                     dump_walkers=2, dump_capacity=900)
    return cfg, extra


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("NDS_FUZZ_CASES", "160"))))
def test_random_configuration_matches_oracle(seed):
    rng = np.random.RandomState(9000 + seed)
    cfg, extra = random_case(rng)
    M, n = 6, 300
    try:
        spec = spec_from(cfg, n_walkers=M, precision="f64", noise_mode=_abi.NDS_NOISE_EXTERNAL, **extra)
    except (ValueError, AssertionError, NotImplementedError, TypeError, NameError) as e:
        pytest.skip(f"configuration refused by the front-end like by the reference: {e}")
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    nd = spec.ndim
    nn = {"langevin": nd, "2nd-langevin": 2 * nd}.get(spec.integrate, 0)
    noise = rng.normal(size=(n, max(nn, 1), M))
    x0 = np.asarray(spec.x0, dtype=float).reshape(-1, 1) + rng.normal(scale=0.3, size=(nd, M))
    v0 = rng.normal(scale=0.01, size=(nd, M))
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        if nn:
            sim.set_noise(noise)
        sim.begin()
    eng.run(170)
    eng.run(n - 170)
    orc.run(n)
    a, b = eng.get_state(), orc.get_state()
    assert np.isfinite(b["x"]).all(), cfg
    for k in ("x", "v", "f", "fixf", "colv", "pe", "ke", "biase", "totale", "T"):
        assert scaled_err(a[k], b[k]) < 1e-8, (k, cfg, extra)
    if "mtd" in cfg["biases"]:
        for wk in (0, M - 1):
            (ce, we), (co, wo) = eng.get_hills(wk), orc.get_hills(wk)
            assert len(we) == len(wo) and np.allclose(we, wo, rtol=1e-8) and np.allclose(ce, co, rtol=1e-8), (cfg, extra)
            if cfg.get("mtd_adapsig"):
                assert np.allclose(eng.get_hill_sigma2(wk), orc.get_hill_sigma2(wk), rtol=1e-7), (cfg, extra)
    if extra.get("dump"):
        fr = eng.drain_frames()
        names = eng.frame_fields()
        assert int(fr[-1, names.index("step"), 0]) == n
        assert scaled_err(fr[-1, names.index("x0"), :], b["x"][0, :2]) < 1e-8
        assert scaled_err(fr[-1, names.index("pe"), :], b["pe"][:2]) < 1e-8


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("NDS_FUZZ_CASES", "60"))))
def test_random_committor_configuration_matches_oracle(seed):
    """Committor shooting with random basin boxes around the start region (1-3 basins, random diffusion time), every
    integrator that draws noise or not, with and without lane compaction: basin ids and exit steps identical."""
    rng = np.random.RandomState(7000 + seed)
    nd = int(rng.choice([2, 5]))
    cv = "Original" if nd == 2 else str(rng.choice(["Proj5dto2d", "Proj5dto2dv2"]))
    nb = int(rng.choice([1, 2, 3]))
    centre = np.array([22.0, 24.0]) if nd == 2 else np.array([14.0, 20.0])
    crit = []
    for _ in range(nb):
        lo = centre + rng.uniform(-6, 4, 2)
        crit.append([[float(lo[0]), float(lo[0] + rng.uniform(0.5, 4))], [float(lo[1]), float(lo[1] + rng.uniform(0.5, 4))]])
    cfg = dict(ndim=nd, mass=1.0, potential="Mueller2d" if nd == 2 else "Mueller5d", colvar_name=cv,
               x0=[22.0, 24.0] if nd == 2 else [14.0, 2.0, 20.0, 1.0, 12.0], method="committor", n_basins=nb, criteria=crit,
               diffusion_time=int(rng.choice([0, 5, 40])), integrate=str(rng.choice(["langevin", "2nd-langevin", "nve"])),
               gamma=float(rng.choice([0.01, 0.1])), temperature=float(rng.choice([300, 900])), dt=1.0, steps=600)
    if nd == 5:
        cfg["true_colvar_name"] = cv
    M, n = 40, 600
    spec = spec_from(cfg, n_walkers=M, precision="f64", noise_mode=_abi.NDS_NOISE_EXTERNAL,
                     steps_per_launch=int(rng.choice([50, 600])))
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    nn = {"langevin": nd, "2nd-langevin": 2 * nd}.get(spec.integrate, 0)
    noise = rng.normal(size=(n, max(nn, 1), M))
    x0 = np.asarray(spec.x0, dtype=float).reshape(-1, 1) + rng.normal(scale=0.5, size=(nd, M))
    v0 = rng.normal(scale=0.02, size=(nd, M))
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        if nn:
            sim.set_noise(noise)
        sim.begin()
        sim.run(n)
    (be, se), (bo, so) = eng.get_commit(), orc.get_commit()
    assert np.array_equal(be, bo) and np.array_equal(se, so), cfg
    assert eng.count_alive() == int((be < 0).sum())
    a, b = eng.get_state(), orc.get_state()
    assert scaled_err(a["x"], b["x"]) < 1e-8, cfg


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("NDS_FUZZ_CASES", "60"))))
def test_random_monte_carlo_configuration_matches_oracle(seed):
    """Metropolis / Wang-Landau with random proposal mode, boundary condition and biases (metadynamics on the
    accepted-move clock included), uniforms supplied by the caller: accept counts, states, hills, histograms."""
    rng = np.random.RandomState(5000 + seed)
    nd = int(rng.choice([1, 2, 5]))
    pot = {1: "DoubleWell1d", 2: str(rng.choice(["Mueller2d", "DoubleWell"])), 5: "Mueller5d"}[nd]
    method = str(rng.choice(["mc", "wl-mc"]))
    allp = bool(rng.rand() < 0.5)
    cfg = dict(ndim=nd, mass=1.0, potential=pot, x0=28.0 if nd == 1 else ([22.0, 24.0] if nd == 2 else [14.0, 2.0, 20.0, 1.0, 12.0]),
               method=method, bounds=[float(rng.choice([0.5, 2.0]))] * nd, global_bounds=[48.0] * nd,
               temperature=float(rng.choice([300, 900])), dt=0.5, steps=400, propose_dim="all" if allp else "single",
               bound_cond=str(rng.choice(["periodic", "no_bound"])))
    if nd == 5:
        cfg.update(colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d")
    cd = 2 if nd != 1 else 1
    biases = [b for b in ("umb", "mtd") if rng.rand() < 0.4]
    if "umb" in biases:
        cfg.update(umb_k=0.05, umb_r0=[float(v) for v in rng.uniform(15, 28, cd)])
    if "mtd" in biases:
        cfg.update(mtd_w=0.02, mtd_sigma=[1.5] * cd, mtd_dep_freq=float(rng.choice([5, 10])), mtd_biasf=8.0, mtd_shared=False)
    cfg["biases"] = biases
    M, n = 16, 400
    spec = spec_from(cfg, n_walkers=M, precision="f64", noise_mode=_abi.NDS_NOISE_EXTERNAL,
                     steps_per_launch=int(rng.choice([1, 33, 400])))
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    u = rng.uniform(size=(n, nd + 1 if allp else 3, M))
    x0 = np.asarray(spec.x0, dtype=float).reshape(-1, 1) + rng.normal(scale=0.3, size=(nd, M))
    v0 = rng.normal(scale=0.02, size=(nd, M))  # begin() sets ke / totale from it; they stay until the first accepted move
    for sim in (eng, orc):
        sim.set_state(x0, v0)
        sim.set_noise(u)
        sim.begin()
    eng.run(150)
    eng.run(n - 150)
    orc.run(n)
    assert np.array_equal(eng.get_mc(), orc.get_mc()), cfg
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "colv", "pe", "biase", "totale"):
        assert scaled_err(a[k], b[k]) < 1e-9, (k, cfg)
    if "mtd" in biases:
        for wk in (0, M - 1):
            (ce, we), (co, wo) = eng.get_hills(wk), orc.get_hills(wk)
            assert len(we) == len(wo) and np.allclose(we, wo, rtol=1e-9) and np.allclose(ce, co, rtol=1e-9), cfg
    if method == "wl-mc":
        assert np.array_equal(eng.get_wl_histogram(0), orc.get_wl_histogram(0)), cfg


HOT = {
    "c2": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], integrate="langevin", gamma=0.1),
    "c2_l2": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], integrate="2nd-langevin", gamma=0.01,
                  biases=["umb"], umb_k=0.1, umb_r0=[22, 24]),
    "c3": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[21.0, 17.0], integrate="langevin", gamma=0.1,
               method="committor", n_basins=2, criteria=[[[20, 25], [30, 35]], [[38, 50], [2, 10]]]),
    "c4": dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
               x0=[18.0, 0.0, 18.0, 0.0, 100.0], integrate="2nd-langevin", md_gamma=0.01, biases=["umb"], umb_k=0.1,
               umb_r0=[18, 18]),
    "c5": dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
               x0=[14.0, 2.0, 20.0, 1.0, 12.0], integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05,
               mtd_sigma=[2.0, 2.0], mtd_dep_freq=20, mtd_biasf=10.0, mtd_shared=True),
    "c5_grid": dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
                    x0=[14.0, 2.0, 20.0, 1.0, 12.0], integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05,
                    mtd_sigma=[2.0, 2.0], mtd_dep_freq=20, mtd_biasf=10.0, mtd_shared=True, mtd_grid=True,
                    mtd_grid_lo=[-8.0, -8.0], mtd_grid_hi=[48.0, 48.0]),
    "1d": dict(ndim=1, mass=1.0, potential="DoubleWell1d", x0=30.0, integrate="langevin", gamma=0.1),
}


@pytest.mark.parametrize("name", sorted(HOT))
@pytest.mark.parametrize("M", [256, 255])
def test_fp32_hot_variants_follow_the_fp64_engine(name, M):
    """Every fp32 kernel family (packed / scalar hot variants, mtd_block, md_block_wide, generic) against the fp64
    engine of the same configuration under the same Philox seed: Langevin dynamics contract, the two streams differ by
    < 2e-3 per normal (21-bit vs 53-bit fields), so 80 steps later the ensembles still coincide to 1e-2 of a length unit
    -- a wrong constant, sign or stream index in one fp32 variant would not."""
    cfg = dict(HOT[name], temperature=300, dt=1.0, seed=123, steps=80)
    rng = np.random.RandomState(4)
    nd = cfg["ndim"]
    x0 = np.asarray(cfg["x0"], dtype=float).reshape(-1, 1) + rng.normal(scale=0.3, size=(nd, M))
    out = {}
    for prec in ("f32", "f64"):
        eng = Engine(spec_from(cfg, n_walkers=M, precision=prec, steps_per_launch=32).to_struct())
        eng.set_state(x0)
        eng.begin()
        eng.run(80)
        out[prec] = (eng.get_state(), eng.kernel_name)
    (a, ka), (b, kb) = out["f32"], out["f64"]
    assert ka != kb
    if M % 2 == 0 and name in ("c2", "c3"):
        assert "f32x2" in ka, ka
    assert np.max(np.abs(a["x"] - b["x"])) < 1e-2, (name, ka, np.max(np.abs(a["x"] - b["x"])))
    assert np.max(np.abs(a["v"] - b["v"])) < 1e-3, (name, ka)
    assert scaled_err(a["pe"], b["pe"]) < 5e-3, (name, ka)
