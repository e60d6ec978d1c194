"""The N-D harmonic / quartic well named by BASELINE.json's north star.  The reference has no such potential
class, so it is an EXTENSION (`potential: HarmonicQuartic`) and its parity is pinned analytically:
for a harmonic well velocity Verlet is a linear map whose n-th power is known in closed form."""
import numpy as np
import pytest
from scipy import stats

from ndsimulator_b200 import _abi, config
from oracle import oracle
from tests.helpers import spec_from, scaled_err

M_AU = 104.23315626367025
kB = 8.6173303e-5


def exact_nve(x0, v0, k, mass, dt, n):
    """n velocity-Verlet steps (md.py:164-198 without noise) of x'' = -(k/m) x, by matrix power."""
    x0, v0, k = (np.asarray(a, dtype=float) for a in (x0, v0, k))
    w2 = k / (M_AU * mass)
    h = w2 * dt * dt / 2
    out_x, out_v = np.empty_like(x0), np.empty_like(v0)
    for d in range(x0.shape[0]):
        A = np.array([[1 - h[d], dt], [-w2[d] * dt * (1 - h[d] / 2), 1 - h[d]]])
        P = np.linalg.matrix_power(A, n)
        out_x[d] = P[0, 0] * x0[d] + P[0, 1] * v0[d]
        out_v[d] = P[1, 0] * x0[d] + P[1, 1] * v0[d]
    return out_x, out_v


def cfg_for(nd, **kw):
    base = dict(ndim=nd, potential="HarmonicQuartic", potential_k=[0.02 * (d + 1) for d in range(nd)],
                potential_center=[1.5 * d for d in range(nd)], x0=[1.0 + 0.5 * d for d in range(nd)], mass=1.0,
                method="md", integrate="nve", temperature=300, dt=1.0, seed=1)
    base.update(kw)
    return base


def test_config_and_oracle_harmonic_nve_is_the_exact_map():
    spec = spec_from(cfg_for(2))
    st = spec.to_struct()
    assert st.potential == _abi.POTENTIALS["HarmonicQuartic"][0]
    assert list(st.pot_params[0:2]) == [0.02, 0.04] and list(st.pot_params[10:12]) == [0.0, 1.5]
    with pytest.raises(ValueError):
        config.parse(cfg_for(2, potential_k=[1.0, 2.0, 3.0]))
    o = oracle.OracleSim(st, seeds=[0])
    x0 = np.array([[1.0], [1.5]])
    v0 = np.array([[0.01], [-0.02]])
    o.set_state(x0, v0)
    o.begin()
    o.run(10000)
    got = o.get_state()
    c = np.array([0.0, 1.5])
    ex, ev = exact_nve(x0[:, 0] - c, v0[:, 0], [0.02, 0.04], 1.0, 1.0, 10000)
    assert scaled_err(got["x"][:, 0] - c, ex) < 1e-10 and scaled_err(got["v"][:, 0], ev) < 1e-10
    u = got["x"][:, 0] - c
    assert got["pe"][0] == pytest.approx(0.5 * (0.02 * u[0] ** 2 + 0.04 * u[1] ** 2), rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("nd", [1, 2, 5])
def test_gpu_harmonic_nve_matches_the_exact_map(nd):
    """fp64, 10^4 steps, 1e-10 relative (the north star's bar for deterministic paths), many start points."""
    from ndsimulator_b200.engine import Engine
    M = 257
    spec = spec_from(cfg_for(nd), n_walkers=M, steps_per_launch=1000)
    eng = Engine(spec.to_struct())
    rng = np.random.RandomState(nd)
    c = np.array([1.5 * d for d in range(nd)])
    k = np.array([0.02 * (d + 1) for d in range(nd)])
    x0 = c.reshape(-1, 1) + rng.normal(scale=2.0, size=(nd, M))
    v0 = rng.normal(scale=0.02, size=(nd, M))
    eng.set_state(x0, v0)
    eng.begin()
    eng.run(10000)
    st = eng.get_state()
    for w in (0, 100, M - 1):
        ex, ev = exact_nve(x0[:, w] - c, v0[:, w], k, 1.0, 1.0, 10000)
        assert scaled_err(st["x"][:, w] - c, ex) < 1e-10, w
        assert scaled_err(st["v"][:, w], ev) < 1e-10, w
    e0 = 0.5 * (k.reshape(-1, 1) * (x0 - c.reshape(-1, 1)) ** 2).sum(0) + 0.5 * M_AU * (v0 ** 2).sum(0)
    assert np.abs(st["totale"] - e0).max() < 2e-3 * np.abs(e0).max()      # velocity Verlet: bounded energy error


@pytest.mark.gpu
def test_gpu_quartic_langevin_external_noise_matches_oracle():
    from ndsimulator_b200.engine import Engine
    M, n = 16, 2000
    cfg = cfg_for(5, integrate="2nd-langevin", gamma=0.05, potential_q=[0.001 * (d + 1) for d in range(5)],
                  colvar_name="Proj5dto2d", biases=["umb"], umb_k=0.05, umb_r0=[2.0, 4.0])
    spec = spec_from(cfg, n_walkers=M, noise_mode=_abi.NDS_NOISE_EXTERNAL, steps_per_launch=300)
    st = spec.to_struct()
    eng, orc = Engine(st), oracle.OracleSim(st, seeds=np.arange(M))
    noise = np.random.RandomState(8).normal(size=(n, 10, M))
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1) + np.random.RandomState(9).normal(scale=0.3, size=(5, M))
    for sim in (eng, orc):
        sim.set_state(x0, np.zeros((5, M)))
        sim.set_noise(noise)
        sim.begin()
        sim.run(n)
    a, b = eng.get_state(), orc.get_state()
    for key in ("x", "v", "f", "fixf", "colv", "pe", "biase", "totale"):
        assert scaled_err(a[key], b[key]) < 1e-9, key


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_harmonic_langevin_stationary_distribution(precision):
    """Philox-driven Langevin in a 2-D harmonic well: positions are Gaussian with variance kB T_eff / k, where
    T_eff carries the reference's same-xi-in-both-half-kicks factor; checked against an oracle
    ensemble (KS) and against the analytic width within 3 %."""
    from ndsimulator_b200.engine import Engine
    M, n = 1 << 14, 4000
    cfg = cfg_for(2, integrate="langevin", gamma=0.1, potential_k=[0.02, 0.05], potential_center=[0.0, 0.0],
                  x0=[0.0, 0.0], seed=5)
    spec = spec_from(cfg, n_walkers=M, precision=precision)
    eng = Engine(spec.to_struct())
    eng.set_state(np.zeros((2, M)))
    eng.begin()
    eng.run(n)
    x = eng.get_state()["x"]
    Mo = 2048
    orc = oracle.OracleSim(spec_from(cfg, n_walkers=Mo).to_struct(), seeds=np.arange(Mo) + 77)
    orc.set_threads(oracle.max_threads())
    orc.set_state(np.zeros((2, Mo)))
    orc.begin()
    orc.run(n)
    xo = orc.get_state()["x"]
    for d, k in enumerate((0.02, 0.05)):
        assert stats.ks_2samp(x[d], xo[d]).pvalue > 1e-3
        c1 = np.exp(-0.05)
        t_eff = 300 * (1 + c1) ** 2 / (1 + c1 ** 2)
        assert x[d].var() == pytest.approx(kB * t_eff / k, rel=0.04)
