"""Metropolis Monte Carlo (reference engines/mc.py, "next" row of SURVEY.md section 8f).

The reference proposes moves with the unseeded GLOBAL numpy RNG, so its runs are not
reproducible: parity is (i) deterministic against the oracle when the uniforms are supplied by the caller,
(ii) statistical (Boltzmann distribution, acceptance ratio) otherwise."""
import numpy as np
import pytest
from scipy import stats

from ndsimulator_b200 import _abi, config
from oracle import oracle
from tests.helpers import spec_from, scaled_err

kB = 8.6173303e-5
DW = dict(ndim=1, potential="DoubleWell1d", x0=30.0, method="mc", mc_bounds=[2.0], mc_global_bounds=[40.0],
          mass=1.0, temperature=3000, seed=0)
M2 = dict(ndim=2, potential="Mueller2d", x0=[22.0, 24.0], method="mc", mc_bounds=[1.5, 1.5],
          mc_global_bounds=[48.0, 40.0], mass=1.0, temperature=600, seed=0)


def test_config_mc():
    s = config.parse(dict(M2, propose_dim="all", bound_cond="no_bound"))
    st = s.to_struct()
    assert st.method == _abi.NDS_METHOD_MC and st.mc_propose_all == 1 and st.mc_periodic == 0
    assert not s.dump
    with pytest.raises(NameError):
        config.parse(dict(M2, mc_bounds=[1.0]))  # mc.py:31-32
    m = config.parse(dict(M2, biases=["mtd"], mtd_w=0.1, mtd_sigma=[1.0, 1.0], mtd_biasf=5.0, steps=100)).to_struct()
    assert m.mtd_enabled == 1 and m.method == _abi.NDS_METHOD_MC
    wl = config.parse(dict(M2, method="wl-mc", biases=["mtd"], mtd_w=0.1, mtd_sigma=[1.0, 1.0], mtd_biasf=5.0))
    assert wl.to_struct().method == _abi.NDS_METHOD_WL and wl.mtd_shared is False  # per-walker deposition clocks


def test_oracle_mc_samples_boltzmann():
    M, n = 3000, 3000
    spec = spec_from(DW, n_walkers=M)
    o = oracle.OracleSim(spec.to_struct(), seeds=np.arange(M))
    o.set_threads(oracle.max_threads())
    o.set_state(np.full((1, M), 30.0))
    o.begin()
    o.run(n)
    x = o.get_state()["x"][0]
    g = np.linspace(0, 40, 4001)[:-1]
    V = -0.5 * (np.exp(-(g - 10) ** 2 / 100) + np.exp(-(g - 30) ** 2 / 100))
    p = np.exp(-V / (kB * 3000))
    cdf = np.cumsum(p) / p.sum()
    assert stats.kstest(x, lambda t: np.interp(t, g + 0.005, cdf)).pvalue > 1e-3
    assert np.all((x >= 0) & (x < 40))  # periodic wrap like numpy %


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["single_2d", "all_2d_umb", "single_5d"])
def test_gpu_mc_external_uniforms_match_oracle(case):
    """With caller-supplied uniforms every accept / reject decision is deterministic: positions, energies
    and acceptance counts of M walkers must equal the oracle's."""
    from ndsimulator_b200.engine import Engine
    M, n = 64, 600
    if case == "single_2d":
        cfg, nu = dict(M2), 3
    elif case == "all_2d_umb":
        cfg, nu = dict(M2, propose_dim="all", biases=["umb"], umb_k=0.05, umb_r0=[25, 20]), 3
    else:
        cfg, nu = dict(ndim=5, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
                       x0=[12.0] * 5, method="mc", mc_bounds=[2.0] * 5, mc_global_bounds=[40.0] * 5, mass=1.0,
                       temperature=300, seed=0), 3
    spec = spec_from(cfg, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=250)
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    assert eng.kernel_name.startswith("mc_block")
    u = np.random.RandomState(7).uniform(size=(n, nu, M))
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    for sim in (eng, orc):
        sim.set_state(x0)
        sim.set_noise(u)
        sim.begin()
        sim.run(n)
    a, b = eng.get_state(), orc.get_state()
    assert np.array_equal(eng.get_mc(), orc.get_mc())
    assert 0.2 < eng.get_mc().mean() / n < 1.0
    for k in ("x", "colv", "pe", "biase"):
        assert scaled_err(a[k], b[k]) < 1e-10, k


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_mc_boltzmann_statistics(precision):
    """Philox-driven MC: stationary distribution = Boltzmann (KS against the exact 1-D density), and the
    acceptance ratio agrees with an oracle ensemble within 5 standard errors."""
    from ndsimulator_b200.engine import Engine
    M, n = 1 << 15, 3000
    spec = spec_from(DW, n_walkers=M, precision=precision)
    eng = Engine(spec.to_struct())
    eng.set_state(np.full((1, M), 30.0))
    eng.begin()
    eng.run(n)
    x = eng.get_state()["x"][0]
    g = np.linspace(0, 40, 4001)[:-1]
    V = -0.5 * (np.exp(-(g - 10) ** 2 / 100) + np.exp(-(g - 30) ** 2 / 100))
    p = np.exp(-V / (kB * 3000))
    cdf = np.cumsum(p) / p.sum()
    assert stats.kstest(x, lambda t: np.interp(t, g + 0.005, cdf)).pvalue > 1e-3
    Mo = 1024
    orc = oracle.OracleSim(spec_from(DW, n_walkers=Mo).to_struct(), seeds=np.arange(Mo))
    orc.set_threads(oracle.max_threads())
    orc.set_state(np.full((1, Mo), 30.0))
    orc.begin()
    orc.run(n)
    ra, rb = eng.get_mc() / n, orc.get_mc() / n
    assert abs(ra.mean() - rb.mean()) < 5 * np.sqrt(ra.var() / M + rb.var() / Mo)


@pytest.mark.gpu
def test_ndrun_mc_frontend(tmp_path):
    from ndsimulator_b200 import NDRun
    sim = NDRun(**dict(M2, steps=500, n_walkers=256, root=str(tmp_path), run_name="mc", screen=False))
    sim.begin()
    sim.run()
    sim.end()
    assert sim.accepted.shape == (256,) and 0 < sim.accept_rate <= 500
    assert sim.time == sim.accept_rate * 0.5  # dt default 0.5, advanced on accepted moves only
    assert "!! MC acceptance ratio" not in open(str(tmp_path / "mc" / "log")).read()  # screen: false


# --------------------------------------------------------------------------- Wang-Landau (wl-mc)
WL = dict(ndim=2, potential="Mueller2d", x0=[22.0, 30.0], method="wl-mc", bounds=[2.0, 2.0],
          global_bounds=[48.0, 40.0], mass=1.0, temperature=300, seed=0)


def test_oracle_wang_landau_flattens_visits():
    """WangLandau.update (engines/wang_landau.py:65-161): visited bins grow by factors of f = 2, so the
    walk is pushed out of the minimum it starts in and covers a wide energy range."""
    M, n = 64, 4000
    spec = spec_from(WL, n_walkers=M)
    o = oracle.OracleSim(spec.to_struct(), seeds=np.arange(M))
    o.set_state(np.repeat(spec.x0.reshape(-1, 1), M, axis=1))
    o.begin()
    o.run(n)
    h = o.get_wl_histogram(0)
    assert (h > 0).sum() > 30                      # many energy bins visited
    assert np.all((h == 0) | (np.log2(np.where(h > 0, h, 1)) % 1 == 0))  # 1, 2, 4, ... (f = 2)
    pe = o.get_state()["pe"]
    assert pe.max() <= 0.2 + 1e-12                 # moves above Emax are never accepted


@pytest.mark.gpu
def test_gpu_wang_landau_external_uniforms_match_oracle():
    from ndsimulator_b200.engine import Engine
    M, n = 48, 1500
    spec = spec_from(WL, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=400)
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    assert "wang-landau" in eng.kernel_name
    u = np.random.RandomState(11).uniform(size=(n, 3, M))
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    for sim in (eng, orc):
        sim.set_state(x0)
        sim.set_noise(u)
        sim.begin()
        sim.run(n)
    assert np.array_equal(eng.get_mc(), orc.get_mc())
    for wk in (0, 7, M - 1):
        assert np.array_equal(eng.get_wl_histogram(wk), orc.get_wl_histogram(wk))
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "pe", "f"):
        assert scaled_err(a[k], b[k]) < 1e-10, k


@pytest.mark.gpu
def test_gpu_wang_landau_histogram_roundtrip_continues_the_run():
    """dump_hE / copy_hE with one basin (wang_landau.py:24-47): a run that is stopped, whose histogram and
    positions are installed on a fresh engine, continues exactly like the uninterrupted run."""
    from ndsimulator_b200.engine import Engine
    M, n, cut = 16, 600, 250
    spec = spec_from(WL, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=100)
    st = spec.to_struct()
    u = np.random.RandomState(12).uniform(size=(n, 3, M))
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    full = Engine(st)
    full.set_state(x0); full.set_noise(u); full.begin(); full.run(n)
    first = Engine(st)
    first.set_state(x0); first.set_noise(u); first.begin(); first.run(cut)
    hists = [first.get_wl_histogram(w) for w in range(M)]
    second = Engine(st)
    second.set_state(first.get_state()["x"])
    second.set_noise(u[cut:])
    second.begin()
    for w in range(M):
        second.set_wl_histogram(hists[w], walker=w)
    second.run(n - cut)
    for w in (0, 5, M - 1):
        assert np.array_equal(second.get_wl_histogram(w), full.get_wl_histogram(w))
    assert np.array_equal(second.get_state()["x"], full.get_state()["x"])
    assert np.array_equal(first.get_mc() + second.get_mc(), full.get_mc())


@pytest.mark.gpu
def test_gpu_wang_landau_philox_statistics():
    """Philox-driven WL walkers explore like the oracle's: distribution of visited-bin counts (KS)."""
    from ndsimulator_b200.engine import Engine
    M, n = 2048, 3000
    spec = spec_from(WL, n_walkers=M)
    eng = Engine(spec.to_struct())
    eng.set_state(np.repeat(spec.x0.reshape(-1, 1), M, axis=1))
    eng.begin()
    eng.run(n)
    Mo = 512
    orc = oracle.OracleSim(spec_from(WL, n_walkers=Mo).to_struct(), seeds=np.arange(Mo))
    orc.set_threads(oracle.max_threads())
    orc.set_state(np.repeat(spec.x0.reshape(-1, 1), Mo, axis=1))
    orc.begin()
    orc.run(n)
    assert stats.ks_2samp(eng.get_mc(), orc.get_mc()).pvalue > 1e-3
    assert stats.ks_2samp(eng.get_state()["pe"], orc.get_state()["pe"]).pvalue > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("adaptive", [False, True])
def test_gpu_mc_metadynamics_matches_oracle(adaptive):
    """Metadynamics under Metropolis MC.  NDRun.time advances on ACCEPTED moves only (ndrun.py:234-237), so
    fix.update (gaus.py:130-146) follows each walker's own clock: walkers deposit at different moves into their
    own tables, inside the kernel.  VR is refreshed by the deposition only (MC passes col0 to fix.energy,
    gaus.py:361-362).  Deterministic against the oracle with caller-supplied uniforms: accept counts, clocks,
    per-walker hill counts and tables (with per-hill variances for adapsig_geo), positions, energies, VR."""
    from ndsimulator_b200.engine import Engine
    M, n = 32, 600
    cfg = dict(M2, dt=0.5, biases=["mtd"], mtd_w=0.02, mtd_dep_freq=10, mtd_biasf=8.0, steps=n)
    if adaptive:
        cfg.update(colvar_name="XmY", mtd_sigma=1.5, mtd_adapsig=True, mtd_adapsig_geo=True)
    else:
        cfg.update(mtd_sigma=[1.5, 1.5])
    spec = spec_from(cfg, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=64)
    assert spec.mtd_shared is False
    with pytest.raises(ValueError):
        config.parse(dict(cfg, mtd_shared=True))
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    assert eng.kernel_name.startswith("mc_block")
    u = np.random.RandomState(3).uniform(size=(n, 3, M))
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    for sim in (eng, orc):
        sim.set_state(x0)
        sim.set_noise(u)
        sim.begin()
        sim.run(n)
    acc = eng.get_mc()
    assert np.array_equal(acc, orc.get_mc())
    assert eng.time == orc.time == acc[0] * 0.5                       # walker 0's clock
    a, b = eng.get_state(), orc.get_state()
    for k in ("x", "colv", "pe", "biase", "f", "fixf", "ke", "totale"):
        assert scaled_err(a[k], b[k]) < 1e-10, k
    assert scaled_err(eng.get_vr(), orc.get_vr()) < 1e-10
    counts = []
    for wk in (0, 7, M - 1):
        (cg, hg), (co, ho) = eng.get_hills(wk), orc.get_hills(wk)
        # the walker's own clock: one deposition per dep_freq / dt = 20 ACCEPTED moves
        assert len(hg) == len(ho) == eng.n_hills(wk)
        lo_n, hi_n = int(np.floor(max(acc[wk] - 1, 0) * 0.5 / 10)) + 1, int(np.floor(acc[wk] * 0.5 / 10)) + 1
        assert lo_n <= len(hg) <= hi_n
        counts.append(len(hg))
        assert scaled_err(cg, co) < 1e-10 and scaled_err(hg, ho) < 1e-10
        assert hg.min() < 0.9 * 0.02                                  # the tempering is active
        if adaptive:
            sg, so = eng.get_hill_sigma2(wk), orc.get_hill_sigma2(wk)
            assert np.array_equal(sg, so) and sg[0] == 2.0 and np.all(sg[1:] == 2.0 * 1.5 ** 2)
    assert len(set(int(c) for c in np.floor(np.maximum(acc - 1, 0) * 0.5 / 10))) > 1   # clocks really differ
    assert b["biase"].max() > 0.01


# ----------------------------------------------------- pinned to the reference's own (seeded) Metropolis runs
MC_GOLD = None


def _mc_gold():
    global MC_GOLD
    if MC_GOLD is None:
        from tests.helpers import load_golden
        MC_GOLD = load_golden("mc.json")
    return MC_GOLD


def _reference_uniforms(rec):
    """The numbers the reference consumed (oracle/make_golden.py mc_runs): proposals from the GLOBAL numpy RNG
    seeded with rec['global_seed'] (mc.py:52-55), acceptance draws from RandomState(seed) (mc.py:82) after the ndim
    velocity draws of begin() (modify.py:44-54)."""
    cfg = rec["config"]
    nd, n = cfg["ndim"], cfg["steps"]
    g = np.random.RandomState(rec["global_seed"])
    rs = np.random.RandomState(cfg["seed"])
    for _ in range(nd):
        rs.normal(0, 1, 1)
    allp = cfg.get("mc_propose_dim", cfg.get("propose_dim", "single")) == "all"
    u = np.zeros((n, nd + 1 if allp else 3, 1))
    for k in range(n):
        if allp:
            u[k, :nd, 0] = g.rand(nd)
            u[k, nd, 0] = rs.rand()
        else:
            dim = g.randint(nd)
            u[k, 0, 0] = (dim + 0.5) / nd          # floor(u * ndim) == dim
            u[k, 1, 0] = g.rand()
            u[k, 2, 0] = rs.rand()
    return u


def _check_against_reference(sim, rec, every=1):
    n = rec["config"]["steps"]
    xs, pes, bes, accs = [], [], [], []
    for k in range(0, n, every):
        sim.run(every)
        st = sim.get_state()
        xs.append(st["x"][:, 0].copy()); pes.append(st["pe"][0]); bes.append(st["biase"][0])
        accs.append(int(sim.get_mc()[0]))
    idx = list(range(every - 1, n, every))
    assert accs == [rec["accepted"][i] for i in idx]
    assert scaled_err(np.array(xs), np.array(rec["x"])[idx]) < 1e-10
    assert scaled_err(np.array(pes), np.array(rec["pe"])[idx]) < 1e-10
    assert scaled_err(np.array(bes), np.array(rec["biase"])[idx]) < 1e-10
    assert sim.time == rec["time"][-1]
    if "hills_heights" in rec:
        c, h = sim.get_hills(0)
        assert len(h) == len(rec["hills_heights"])
        assert scaled_err(c, np.array(rec["hills_centers"])) < 1e-10 and scaled_err(h, np.array(rec["hills_heights"])) < 1e-10
        assert sim.get_vr()[0] == pytest.approx(rec["VR"], rel=1e-9)
        if "hills_sigma2" in rec:
            assert np.allclose(sim.get_hill_sigma2(0), rec["hills_sigma2"], rtol=1e-12)
    if "hE" in rec:  # Wang-Landau histogram (wang_landau.py:17,51-55)
        assert np.array_equal(sim.get_wl_histogram(0), np.array(rec["hE"]))


@pytest.mark.gpu
def test_gpu_mc_metadynamics_full_table_fails_loudly():
    """A walker that has a hill to deposit and no room left in its table must not silently stop depositing: nds_run
    fails with the MD path's 'hill table full'."""
    from ndsimulator_b200.engine import Engine, NdsError
    M, n = 8, 600
    cfg = dict(M2, dt=0.5, biases=["mtd"], mtd_w=0.02, mtd_dep_freq=10, mtd_biasf=8.0, mtd_sigma=[1.5, 1.5],
               steps=n, mtd_capacity=5)
    spec = spec_from(cfg, n_walkers=M, steps_per_launch=64)
    eng = Engine(spec.to_struct())
    eng.set_state(np.repeat(spec.x0.reshape(-1, 1), M, axis=1))
    eng.begin()
    eng.run(80)                        # at most 80 accepted moves: floor(80 * 0.5 / 10) + 1 = 5 hills fit
    assert max(eng.n_hills(w) for w in range(M)) <= 5
    with pytest.raises(NdsError, match="hill table full"):
        eng.run(n)


MC_CASES = ["single_2d", "all_2d_umb", "single_2d_mtd", "single_2d_mtd_adapsig_geo", "wl_single_2d", "wl_single_2d_mtd"]


@pytest.mark.parametrize("name", MC_CASES)
def test_oracle_mc_follows_the_seeded_reference(name):
    """Every move of the reference's Metropolis run (accept / reject decisions, positions, energies, NDRun.time that
    only advances on accepted moves, the hills deposited on that clock and their adapsig_geo variances)."""
    rec = _mc_gold()[name]
    spec = spec_from(rec["config"], n_walkers=1, noise_mode=_abi.NDS_NOISE_EXTERNAL)
    o = oracle.OracleSim(spec.to_struct(), seeds=[rec["config"]["seed"]])
    o.set_state(spec.x0.reshape(-1, 1))
    o.set_noise(_reference_uniforms(rec))
    o.begin()
    _check_against_reference(o, rec, every=1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MC_CASES)
def test_gpu_mc_follows_the_seeded_reference(name):
    from ndsimulator_b200.engine import Engine
    rec = _mc_gold()[name]
    spec = spec_from(rec["config"], n_walkers=1, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=64)
    eng = Engine(spec.to_struct())
    eng.set_state(spec.x0.reshape(-1, 1))
    eng.set_noise(_reference_uniforms(rec))
    eng.begin()
    _check_against_reference(eng, rec, every=40)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mc", "mc_mtd", "wl", "wl_mtd", "mc_all_umb"])
def test_gpu_mc_checkpoint_resume_is_exact(case):
    """Metropolis / Wang-Landau ensembles continue bit for bit from a checkpoint on a fresh engine: accepted moves,
    the per-walker clocks and hill tables of MC metadynamics, Wang-Landau histograms, and the event-driven ke / totale
    (mc.py:93-95) come back; the Philox proposal stream is a function of (seed, walker, move index)."""
    from ndsimulator_b200.engine import Engine
    M, n, cut = 48, 600, 250
    cfg = dict(M2, dt=0.5, steps=n, seed=5)
    if case.startswith("wl"):
        cfg.update(method="wl-mc", bounds=[1.5, 1.5], global_bounds=[48.0, 40.0])
        cfg.pop("mc_bounds"), cfg.pop("mc_global_bounds")
    if case.endswith("mtd"):
        cfg.update(biases=["mtd"], mtd_w=0.02, mtd_sigma=[1.5, 1.5], mtd_dep_freq=10, mtd_biasf=8.0)
    if case == "mc_all_umb":
        cfg.update(propose_dim="all", biases=["umb"], umb_k=0.05, umb_r0=[25, 20])
    spec = spec_from(cfg, n_walkers=M, steps_per_launch=64)
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1) + np.random.RandomState(1).normal(scale=0.4, size=(2, M))

    def fresh():
        e = Engine(spec.to_struct())
        e.set_state(x0)
        e.begin()
        return e

    a = fresh()
    a.run(n)
    b = fresh()
    b.run(cut)
    ck = b.checkpoint()
    c = Engine(spec.to_struct())
    c.restore(ck)
    c.run(n - cut)
    assert c.step == a.step == n
    assert np.array_equal(c.get_mc(), a.get_mc()) and a.get_mc().sum() > 0
    sa, sc = a.get_state(), c.get_state()
    for k in ("x", "colv", "pe", "ke", "biase", "totale"):
        assert np.array_equal(sa[k], sc[k]), k
    if case.endswith("mtd"):
        assert np.array_equal(a.get_vr(), c.get_vr())
        counts = set()
        for wk in (0, 17, M - 1):
            (ca, ha), (cc, hc) = a.get_hills(wk), c.get_hills(wk)
            assert np.array_equal(ha, hc) and np.array_equal(ca, cc) and len(ha) > 3
            counts.add(len(ha))
    if case.startswith("wl"):
        for wk in (0, M - 1):
            assert np.array_equal(a.get_wl_histogram(wk), c.get_wl_histogram(wk))
        assert a.get_wl_histogram(0).sum() > 0
