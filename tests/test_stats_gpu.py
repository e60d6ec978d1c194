"""GPU statistical-parity tests for the stochastic paths (Philox noise cannot reproduce MT19937 draws).

The comparison ensembles come from the oracle, which reproduces the reference's own stochastic runs bit
for bit (tests/test_oracle_golden.py), i.e. "the reference's own runs" of BASELINE.json north_star.
Acceptance thresholds are stated per test; the KS tests use p > 1e-3 on independent walkers.
"""
import numpy as np
import pytest
from scipy import stats

from ndsimulator_b200.engine import Engine
from oracle import oracle
from tests.helpers import load_golden, spec_from, philox_normals_f64

pytestmark = pytest.mark.gpu

MD2D = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md", temperature=300,
            dt=1.0, seed=0)
kB = 8.6173303e-5


def test_initial_velocities_are_philox_normals():
    """v0 = sqrt(kB T / m) * N(0,1) (modify.py:44-54) from the documented Philox stream (domain 1)."""
    M = 1 << 16
    spec = spec_from(dict(MD2D, integrate="nve"), n_walkers=M, walker_offset=5, n_walkers_global=M + 5,
                     seed=1234567890123)
    eng = Engine(spec.to_struct())
    eng.set_state(np.repeat(spec.x0.reshape(-1, 1), M, axis=1))
    eng.begin()
    st = eng.get_state()
    scale = np.sqrt(kB * 300 / (104.23315626367025 * 1.0))
    z = st["v"] / scale
    for w in (0, 1, 777, M - 1):  # exact stream check against the Python restatement
        want = philox_normals_f64(1234567890123, w + 5, 2, domain=1)
        assert np.allclose(z[:, w], want, rtol=1e-12, atol=1e-14)
    assert stats.kstest(z.reshape(-1), "norm").pvalue > 1e-3
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1) < 0.02
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.02
    assert st["T"].mean() == pytest.approx(300.0, rel=0.02)


@pytest.mark.parametrize("case", ["1d", "5d", "2d_langevin2"])
def test_step_noise_is_the_documented_philox_stream(case):
    """Every lane of the six-normals-per-call layout, and the step grouping (1-D: six steps per call, 5-D: one
    call per step with the sixth normal dropped, 2-D 2nd-order Langevin: four normals per step, two calls per
    three steps): on a flat potential the Langevin update is linear in the noise, so the normals the kernel
    used can be solved from consecutive velocities and compared with the Python restatement of the stream
    (tests/helpers.py).  Launch boundaries fall inside groups on purpose."""
    M, n = 64, 13
    if case == "1d":
        cfg, nd, nn, grp = dict(ndim=1, potential="Flat1d", x0=0.0, integrate="langevin"), 1, 1, 6
    elif case == "5d":
        cfg, nd, nn, grp = dict(ndim=5, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
                                x0=[12.0] * 5, integrate="langevin"), 5, 5, 1
    else:
        cfg, nd, nn, grp = dict(ndim=2, potential="Flat2d", x0=[0.0, 0.0], integrate="2nd-langevin"), 2, 4, 3
    cfg = dict(cfg, mass=1.0, method="md", temperature=300, dt=1.0, gamma=0.1, seed=42)
    spec = spec_from(cfg, n_walkers=M, steps_per_launch=4, walker_offset=3, n_walkers_global=M + 3)
    eng = Engine(spec.to_struct())
    x0 = np.repeat(np.asarray(spec.x0, dtype=float).reshape(-1, 1), M, axis=1)
    eng.set_state(x0, np.zeros((nd, M)))
    eng.begin()
    st = eng.get_state()
    vs, fs = [st["v"].copy()], [st["f"].copy()]
    for _ in range(n):
        eng.run(1)
        st = eng.get_state()
        vs.append(st["v"].copy())
        fs.append(st["f"].copy())
    vs, fs = np.array(vs), np.array(fs)                    # [n+1][nd][M]
    m = 104.23315626367025
    calls_per_group = -(-grp * nn // 6)
    for w in (0, 17, M - 1):
        z = philox_normals_f64(42, w + 3, 6 * calls_per_group * (n // grp + 1))
        for s in range(n):
            g, i = divmod(s, grp)
            zs = z[6 * calls_per_group * g + i * nn: 6 * calls_per_group * g + (i + 1) * nn]
            if case != "2d_langevin2":   # v' = c1 (c1 (v + h f) + c2 xi + h f') + c2 xi, h = dt / 2m  (md.py:164-198)
                c1 = np.exp(-0.05)
                c2 = np.sqrt((1 - c1 ** 2) * kB * 300 / m)
                h = 0.5 / m
                got = (vs[s + 1, :, w] - c1 * h * fs[s + 1, :, w] - c1 * c1 * (vs[s, :, w] + h * fs[s, :, w])) / (c2 * (1 + c1))
                assert np.allclose(got, zs[:nd], rtol=1e-7, atol=1e-9), (w, s)
            else:                                          # two half updates v += -c2 v + c3 xi - c4 eta
                frdt = 8 * (1 - np.exp(-0.05))
                sig = np.sqrt(2 * kB * 300 * frdt / m)
                k2 = frdt / 2 - frdt ** 2 / 8
                k3 = sig / 2 * (1 - frdt / 4)
                k5 = sig / (2 * np.sqrt(nd))
                k4 = frdt / 2 * k5
                xi, eta = zs[:nd], zs[nd:]
                v1 = vs[s, :, w] * (1 - k2) + k3 * xi - k4 * eta
                want = v1 * (1 - k2) + k3 * xi - k4 * eta
                assert np.allclose(vs[s + 1, :, w], want, rtol=1e-9, atol=1e-15), (w, s)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_fp32_normals_moments(precision):
    """Per-step noise: on a flat potential one Langevin step from v=0 gives v = c2 (1 + c1) xi."""
    M = 1 << 18
    cfg = dict(MD2D, potential="Flat2d", gamma=0.1, integrate="langevin")
    spec = spec_from(cfg, n_walkers=M, precision=precision)
    eng = Engine(spec.to_struct())
    eng.set_state(np.zeros((2, M)), np.zeros((2, M)))
    eng.begin()
    eng.run(1)
    v = eng.get_state()["v"]
    c1 = np.exp(-0.05)
    c2 = np.sqrt((1 - c1 ** 2) * kB * 300 / 104.23315626367025)
    z = (v / (c2 * (1 + c1))).reshape(-1)
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1) < 0.01
    assert abs(stats.kurtosis(z)) < 0.05
    assert np.abs(z).max() > 4.0  # tails are populated


@pytest.mark.parametrize("integrate,gamma,dt", [("langevin", 0.1, 1.0), ("langevin", 1.0, 1.0),
                                                ("2nd-langevin", 0.01, 1.0), ("2nd-langevin", 0.1, 2.0)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_effective_temperature_quirks(integrate, gamma, dt, precision):
    """The reference's integrators do not sample T: free-particle stationary variance is
    (kBT/m)(1+c1)^2/(1+c1^2) for langevin and (1+a)^2 (c3^2+c4^2)/(1-a^4) for 2nd-langevin.  The engine
    must reproduce those, not the textbook value."""
    M = 1 << 16
    cfg = dict(MD2D, potential="Flat2d", gamma=gamma, dt=dt, integrate=integrate)
    spec = spec_from(cfg, n_walkers=M, precision=precision)
    st = spec.to_struct()
    eng = Engine(st)
    eng.set_state(np.zeros((2, M)))
    eng.begin()
    eng.run(400)
    T = eng.get_state()["T"].mean()
    c = oracle.OracleSim(spec_from(cfg, n_walkers=1).to_struct(), seeds=[0]).constants()
    m = c["m"]
    if integrate == "langevin":
        var = (kB * 300 / m) * (1 + c["c1"]) ** 2 / (1 + c["c1"] ** 2)
    else:
        a = 1 - c["c2"]
        var = (1 + a) ** 2 * (c["c3"] ** 2 + c["c4"] ** 2) / (1 - a ** 4)
    T_eff = m * var / kB
    assert T == pytest.approx(T_eff, rel=0.02)
    assert abs(T_eff - 300) > 1 or integrate == "2nd-langevin"


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_cv_histograms_match_oracle_ensemble(precision):
    """C2 physics (2-D Mueller Langevin, gamma=0.1): two-sample KS on the x and y marginals and on the
    potential energy of independent walkers after 2000 steps, GPU (Philox) vs oracle (MT19937)."""
    M, nsteps = 4096, 2000
    cfg = dict(MD2D, integrate="langevin", gamma=0.1)
    spec = spec_from(cfg, n_walkers=M, precision=precision)
    st = spec.to_struct()
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    eng = Engine(st)
    eng.set_state(x0)
    eng.begin()
    eng.run(nsteps)
    a = eng.get_state()
    orc = oracle.OracleSim(st, seeds=np.arange(1000, 1000 + M))
    orc.set_threads(oracle.max_threads())
    orc.set_state(x0)
    orc.begin()
    orc.run(nsteps)
    b = orc.get_state()
    for name, u, v in (("x", a["x"][0], b["x"][0]), ("y", a["x"][1], b["x"][1]), ("pe", a["pe"], b["pe"]),
                       ("T", a["T"], b["T"])):
        p = stats.ks_2samp(u, v).pvalue
        assert p > 1e-3, (name, p)
    # <T> agrees within the standard error of the two ensembles (and shows the A1 signature ~ 2T)
    se = np.sqrt(a["T"].var() / M + b["T"].var() / M)
    assert abs(a["T"].mean() - b["T"].mean()) < 5 * se
    assert 500 < a["T"].mean() < 700
    # on-device histogram equals a host histogram of the same state
    h = eng.colvar_histogram(0.0, 1.0, 48, 0.0, 1.0, 40)
    hh, _, _ = np.histogram2d(a["colv"][1], a["colv"][0], bins=[np.arange(0, 41.0), np.arange(0, 49.0)])
    if precision == "f64":
        assert np.array_equal(h, hh.astype(np.int64))
    else:
        assert np.abs(h - hh).sum() <= 4  # fp32 -> fp64 bin-edge rounding


def test_committor_fraction_within_binomial_ci():
    """C3: p(basin 1 | committed) from 8192 GPU shots vs 2048 oracle shots (and the reference's own 24
    shots in tests/golden/committor.json): two-proportion z-test at 95 % (z < 1.96 -> use 3 for a
    deterministic CI-friendly bound), exit-step distribution by KS."""
    g = load_golden("committor.json")
    cfg = dict(g["config"])
    Mg, Mo = 8192, 2048
    spec = spec_from(cfg, n_walkers=Mg, precision="f64")
    eng = Engine(spec.to_struct())
    eng.set_state(np.repeat(spec.x0.reshape(-1, 1), Mg, axis=1))
    eng.begin()
    eng.run(cfg["steps"])
    bg, sg = eng.get_commit()
    spec_o = spec_from(cfg, n_walkers=Mo)
    orc = oracle.OracleSim(spec_o.to_struct(), seeds=np.arange(100, 100 + Mo))
    orc.set_threads(oracle.max_threads())
    orc.set_state(np.repeat(spec.x0.reshape(-1, 1), Mo, axis=1))
    orc.begin()
    orc.run(cfg["steps"])
    bo, so = orc.get_commit()

    def frac(b):
        c = (b >= 0).sum()
        return (b == 1).sum() / c, c

    pg, ng = frac(bg)
    po, no = frac(bo)
    p = (pg * ng + po * no) / (ng + no)
    z = abs(pg - po) / np.sqrt(p * (1 - p) * (1 / ng + 1 / no))
    assert z < 3.0, (pg, po, z)
    # committed fraction
    cg, co = (bg >= 0).mean(), (bo >= 0).mean()
    pc = (cg * Mg + co * Mo) / (Mg + Mo)
    assert abs(cg - co) / np.sqrt(pc * (1 - pc) * (1 / Mg + 1 / Mo)) < 3.0
    assert stats.ks_2samp(sg[bg >= 0], so[bo >= 0]).pvalue > 1e-3
    # the reference's own 24 shots: 6 -> basin 0, 8 -> basin 1 (binomial 95 % CI contains pg)
    ref = np.array([s["basin"] for s in g["shots"]])
    k, n = (ref == 1).sum(), (ref >= 0).sum()
    lo, hi = stats.binomtest(int(k), int(n)).proportion_ci(0.95)
    assert lo <= pg <= hi


def test_mtd_fes_matches_oracle_ensemble():
    """Free-energy surface estimate (sum of hills on the 48x40 grid, gaus.py:365-393) of a single-walker
    well-tempered run (the parameters of examples/2d-mtd.yaml), averaged over 8192 independent runs on either side:
    GPU vs oracle within the north star's plain 0.1 kT RMS over the filled region after aligning the additive
    constant.  8192 runs bring the statistical error of either mean below 0.03 kT (one run scatters by 2.3 kT), so
    the bar is not met by noise."""
    R, nsteps = 8192, 20000
    cfg = dict(MD2D, x0=[22.0, 30.0], integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05,
               mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, steps=nsteps, mtd_shared=False)
    spec = spec_from(cfg, n_walkers=R, precision="f64")
    st = spec.to_struct()
    x0 = np.repeat(spec.x0.reshape(-1, 1), R, axis=1)
    eng = Engine(st)
    eng.set_state(x0)
    eng.begin()
    eng.run(nsteps)
    orc = oracle.OracleSim(st, seeds=np.arange(R))
    orc.set_threads(oracle.max_threads())
    orc.set_state(x0)
    orc.begin()
    orc.run(nsteps)
    ga_all = np.array([eng.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40, walker=w) for w in range(R)])
    gb_all = np.array([orc.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40, walker=w) for w in range(R)])
    ga, gb = ga_all.mean(axis=0), gb_all.mean(axis=0)
    kT = kB * 300
    mask = gb > 2 * kT  # region the bias has filled
    assert mask.sum() > 50
    d = (ga - gb)[mask]
    d = d - d.mean()
    rms = np.sqrt((d ** 2).mean())
    sem_o = gb_all[:, mask].std(axis=0).mean() / np.sqrt(R)
    sem_g = ga_all[:, mask].std(axis=0).mean() / np.sqrt(R)
    assert sem_o < 0.035 * kT and sem_g < 0.035 * kT, (sem_o / kT, sem_g / kT)
    assert rms < 0.1 * kT, (rms / kT, sem_o / kT, sem_g / kT)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_six_normals_of_one_call_are_independent_standard_normals(precision):
    """All six lanes of the 21-bit Box-Muller layout over 2^16 walkers: on a flat 1-D potential six Langevin steps
    consume exactly one Philox call per walker and the normals can be solved from the velocities.  Each lane is
    N(0,1) (KS, variance, kurtosis), lanes are uncorrelated (also in their squares), nothing exceeds the 5.40 sigma
    cut-off of a 21-bit radius."""
    M = 1 << 16
    cfg = dict(ndim=1, potential="Flat1d", x0=0.0, integrate="langevin", mass=1.0, method="md", temperature=300,
               dt=1.0, gamma=0.1, seed=2718)
    spec = spec_from(cfg, n_walkers=M, precision=precision)
    eng = Engine(spec.to_struct())
    eng.set_state(np.zeros((1, M)), np.zeros((1, M)))
    eng.begin()
    vs = [np.zeros(M)]
    for _ in range(6):
        eng.run(1)
        vs.append(eng.get_state()["v"][0].copy())
    c1 = np.exp(-0.05)
    c2 = np.sqrt((1 - c1 ** 2) * kB * 300 / 104.23315626367025)
    z = np.array([(vs[s + 1] - c1 * c1 * vs[s]) / (c2 * (1 + c1)) for s in range(6)])      # [6][M]
    for lane in range(6):
        assert stats.kstest(z[lane], "norm").pvalue > 1e-3, lane
        assert abs(z[lane].var() - 1) < 0.02 and abs(stats.kurtosis(z[lane])) < 0.1, lane
    cc = np.corrcoef(z)
    c2m = np.corrcoef(z ** 2)
    off = ~np.eye(6, dtype=bool)
    assert np.abs(cc[off]).max() < 5 / np.sqrt(M) and np.abs(c2m[off]).max() < 5 / np.sqrt(M)
    assert np.abs(z).max() <= np.sqrt(2 * np.log(2.0 ** 21)) + (1e-9 if precision == "f64" else 1e-3)
    # walkers are independent too: neighbouring walkers (the two lanes of a packed thread)
    assert abs(np.corrcoef(z[:, 0::2].ravel(), z[:, 1::2].ravel())[0, 1]) < 5 / np.sqrt(3 * M)


def test_noise_tails_fp32_ten_billion_draws():
    """Tail mass of the fp32 step noise over 1.007e10 normals drawn on the device through normals6<float>: counts of
    |z| > 4, 4.5, 5 follow the exact tail mass of the 21-bit midpoint scheme (tests/helpers.py) within 5 sigma of the
    Poisson error -- which is the Gaussian tail within 0.3 % / 1 % / 3 % -- nothing lies beyond the documented
    cut-off sqrt(2 ln 2^22) = 5.52 sigma, and the second and fourth moments are those of N(0, 1)."""
    from ndsimulator_b200.engine import noise_tail_counts
    from tests.helpers import fp32_tail_probability
    thr = [3.0, 4.0, 4.5, 5.0, 5.4, 5.53]
    counts, st = noise_tail_counts("f32", 987654321, 1 << 20, 1600, thr)
    n = st["n"]
    assert n > 1e10
    for t, c in zip(thr[:5], counts[:5]):
        want = fp32_tail_probability(t) * n
        assert abs(c - want) < 5 * np.sqrt(want) + 1e-4 * want, (t, c, want)
    assert counts[3] == pytest.approx(2 * stats.norm.sf(5.0) * n, rel=0.05)   # the Gaussian itself, 5 sigma
    assert counts[5] == 0 and 5.3 < st["max"] <= np.sqrt(2 * np.log(2.0 ** 22)) * (1 + 1e-5)
    assert abs(st["mean"]) < 5 / np.sqrt(n)
    assert abs(st["m2"] - 1) < 5 * np.sqrt(2 / n) + 2e-6
    assert abs(st["m4"] - 3) < 5 * np.sqrt(96 / n) + 3e-5


def test_noise_tails_fp64_reach_beyond_five_and_a_half_sigma():
    """The fp64 (parity-precision) engine draws 53-bit radius variates: over 2e9 normals the counts of |z| > 4 ... 5.5
    are the Gaussian's own within 5 sigma of the Poisson error, and draws beyond 5.52 sigma (impossible with 21-bit
    fields) do occur."""
    from ndsimulator_b200.engine import noise_tail_counts
    thr = [4.0, 4.5, 5.0, 5.5, 6.0]
    counts, st = noise_tail_counts("f64", 24681357, 1 << 20, 320, thr)
    n = st["n"]
    assert n > 2e9
    for t, c in zip(thr, counts):
        want = 2 * stats.norm.sf(t) * n
        assert abs(c - want) < 5 * np.sqrt(want) + 1, (t, c, want)
    assert counts[3] > 40 and st["max"] > 5.7
    assert abs(st["mean"]) < 5 / np.sqrt(n) and abs(st["m2"] - 1) < 5 * np.sqrt(2 / n) and abs(st["m4"] - 3) < 5 * np.sqrt(96 / n)
