"""GPU: the NDRun-shaped front-end, the reference-format dump files and the CLI."""
import os

import numpy as np
import pytest

from ndsimulator_b200 import NDRun
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def tokens(text):
    return [ln.split() for ln in text.split("\n")]


def assert_same_text(got, want, rtol=1e-9, what=""):
    """Same lines and tokens; numeric tokens agree to rtol, everything else exactly."""
    g, w = tokens(got), tokens(want)
    assert len(g) == len(w), (what, len(g), len(w))
    for i, (a, b) in enumerate(zip(g, w)):
        assert len(a) == len(b), (what, i, a, b)
        for x, y in zip(a, b):
            try:
                fx, fy = float(x.strip("[]")), float(y.strip("[]"))
            except ValueError:
                assert x == y, (what, i, x, y)
                continue
            assert fx == pytest.approx(fy, rel=rtol, abs=1e-13), (what, i, x, y)
            assert x.startswith("[") == y.startswith("["), (what, i, x, y)


@pytest.mark.parametrize("name", ["2d_nve_dump10", "2d_mtd_gamma0_dump100", "5d_umb_nve_stat10", "2d_xmy_nve",
                                  "2d_mtd_adapsig_geo_xmy_gamma0", "2d_md_yaml_langevin_dump10",
                                  "2d_umb_mtd_gamma0_dump20"])
def test_dump_files_match_reference(name, tmp_path):
    """pos/force/velocity/colvar/thermo/fix*.dat against the files the Python reference wrote
    (tests/golden/dump_files.json), including the lagged thermo row (A16).  The stochastic case (examples/2d-md.yaml
    as shipped: Langevin) is fed the Gaussian numbers RandomState(seed) handed to the reference."""
    rec = load_golden("dump_files.json")[name]
    cfg = dict(rec["config"], root=str(tmp_path), run_name="r", screen=False, precision="f64")
    stochastic = name == "2d_md_yaml_langevin_dump10"
    if stochastic:
        from ndsimulator_b200 import _abi
        cfg["noise_mode"] = _abi.NDS_NOISE_EXTERNAL
    sim = NDRun(**cfg)
    sim.modify.set_velocity(v=np.array(rec["v0"]))
    if stochastic:
        rs = np.random.RandomState(cfg["seed"])
        for _ in range(2):
            rs.normal(0, 1, 1)                                  # the velocity draws of begin() (modify.py:44-54)
        sim.engine.set_noise(np.array([rs.normal(0, 1, 2) for _ in range(cfg["steps"])]).reshape(-1, 2, 1))
    sim.begin()
    sim.run()
    sim.end()
    for fn, want in rec["files"].items():
        with open(os.path.join(str(tmp_path), "r", fn)) as f:
            got = f.read()
        assert_same_text(got, want, what=f"{name}/{fn}")
    log = open(os.path.join(str(tmp_path), "r", "log")).read()
    assert f"end condition {cfg['steps']} {cfg['steps']} True" in log


@pytest.mark.parametrize("name,stop", [("2d_mtd_gamma0_dump100", 150), ("2d_mtd_gamma0_dump100", 200),
                                       ("5d_umb_nve_stat10", 30), ("2d_mtd_adapsig_geo_xmy_gamma0", 70),
                                       ("2d_umb_mtd_gamma0_dump20", 70)])
def test_dump_files_after_checkpoint_resume(name, stop, tmp_path):
    """Stop at `stop`, save, load into a fresh NDRun (which appends to the files), finish: the dump files are still
    the reference's -- no duplicate step-0 row, the hill listing of fix{i}.dat and the lagged thermo sample survive."""
    rec = load_golden("dump_files.json")[name]
    cfg = dict(rec["config"], root=str(tmp_path), run_name="r", screen=False, precision="f64")
    sim = NDRun(**cfg)
    sim.modify.set_velocity(v=np.array(rec["v0"]))
    sim.begin()
    sim.run(stop)
    ck = os.path.join(str(tmp_path), "ck.npz")
    sim.save_checkpoint(ck)
    sim.end()
    sim = NDRun(**cfg)
    sim.load_checkpoint(ck)
    assert sim.step == stop and sim.engine.frames_pending() == 0
    sim.run()
    sim.end()
    for fn, want in rec["files"].items():
        with open(os.path.join(str(tmp_path), "r", fn)) as f:
            got = f.read()
        assert_same_text(got, want, what=f"{name}/{fn}")


def test_ndrun_protocol_and_atoms_view(tmp_path):
    cfg = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md", integrate="nve",
               temperature=300, dt=1.0, seed=7, steps=500, dump=False, screen=False, root=str(tmp_path),
               n_walkers=4)
    sim = NDRun(**cfg)
    sim.begin()
    e0 = sim.atoms.totale
    assert sim.atoms.positions.tolist() == [22.0, 24.0]
    sim.run()
    sim.end()
    assert sim.step == 500 and sim.time == 500.0
    assert abs(sim.atoms.totale - e0) < 1e-3          # NVE: bounded energy error (1.9e-4)
    assert sim.positions.shape == (2, 4)
    assert not np.allclose(sim.positions[:, 0], sim.positions[:, 1])  # independent initial velocities
    # set_seed + injected velocities reproduce a run exactly
    sim2 = NDRun(**cfg)
    sim2.begin()
    sim2.run()
    assert np.array_equal(sim2.positions, sim.positions)


def test_committor_frontend(tmp_path):
    g = load_golden("committor.json")
    cfg = dict(g["config"], root=str(tmp_path), run_name="c", dump=False, screen=False, n_walkers=512,
               seed=3, precision="f32")
    sim = NDRun(**cfg)
    sim.begin()
    sim.run()
    sim.end()
    assert sim.commit_basin in (-1, 0, 1)
    assert sim.commit_basins.shape == (512,)
    committed = sim.commit_basins >= 0
    assert 0.2 < committed.mean() < 0.7               # reference: 205/480 committed within 10^4 steps
    assert np.all(sim.exit_steps[~committed] == cfg["steps"])
    log = open(os.path.join(str(tmp_path), "c", "log")).read()
    assert "end condition" in log


def test_cli_runs_reference_yaml(tmp_path):
    import yaml
    from ndsimulator_b200.scripts.run import main
    cfg = dict(root=str(tmp_path), run_name="cli", ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0],
               method="md", integration="langevin", gamma=1, temperature=300, dt=1.0, steps=200, dump=True,
               dump_freq=10, movie=False, oneplot=True, plot=True, plot_boundary=[[0.0, 48.0], [0.0, 40.0]],
               verbose="debug", screen=False)  # examples/2d-md.yaml keys, including the ignored ones
    path = os.path.join(str(tmp_path), "run.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    sim = main([path, "--set", "n_walkers=8", "dump_walkers=2"])
    assert sim.spec.integrate == "langevin"          # `integration:` is not a key
    assert sim.step == 200
    rows = open(os.path.join(str(tmp_path), "cli", "pos.dat")).read().strip().split("\n")
    assert len(rows) == 21
    assert os.path.exists(os.path.join(str(tmp_path), "cli", "walker_1", "thermo.dat"))


def test_user_defined_python_colvar_is_rejected(tmp_path):
    with pytest.raises(ValueError, match="cannot run on the B200 engine"):
        NDRun(ndim=2, potential="Mueller2d", x0=[1.0, 2.0], colvar_name="my_colvar.NewColvar",
              root=str(tmp_path), screen=False)


@pytest.mark.parametrize("name", ["c2-2d-mueller-langevin", "c3-2d-committor", "c4-5d-umbrella-windows",
                                  "c5-5d-multiwalker-mtd", "mc-2d-metropolis-mtd", "harmonic-quartic-5d",
                                  "c2-side-work-frames-histogram", "c5-5d-multiwalker-mtd-grid", "1d-mtd-adaptive-sigma"])
def test_shipped_examples_run(name, tmp_path):
    """examples/*.yaml (the BASELINE configurations) through the CLI, shortened."""
    from ndsimulator_b200.scripts.run import main
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sim = main([os.path.join(root, "examples", name + ".yaml"), "--set", f"root={tmp_path}", "steps=300",
                "n_walkers=4096"])
    assert sim.step == 300 or sim.spec.method == "committor"
    assert sim.engine.count_nonfinite() == 0
    if name.startswith("c4"):
        c = sim.colv
        k, r0 = sim._window
        assert np.abs(c - r0).mean() < 2.0          # walkers stay near their own window
    if name == "c5-5d-multiwalker-mtd":
        assert sim.engine.n_hills() == 3 * 4096 and sim.engine.kernel_name.startswith("mtd_block")
    if name == "c5-5d-multiwalker-mtd-grid":
        assert sim.engine.n_hills() == 3 * 4096 and "mtd-grid" in sim.engine.kernel_name
        assert sim.engine.get_mtd_grid()[..., 0].max() > 0
    if name.startswith("c2-side-work"):
        assert "frames+hist" in sim.engine.kernel_name
        assert sim.engine.get_histogram().sum() > 0.99 * 30 * 4096      # 30 dump steps, every walker counted
        assert os.path.exists(os.path.join(str(tmp_path), "c2_side_work", "walker_7", "pos.dat"))
    if name.startswith("1d-mtd-adaptive"):
        s2 = sim.engine.get_hill_sigma2(0)
        assert len(s2) == 6 and s2[0] == 1.0 and len(np.unique(s2)) == 6  # first hill: one sample -> variance 1
    if name.startswith("mc"):
        assert sim.engine.kernel_name.startswith("mc_block")
        n0 = sim.engine.n_hills(0)
        assert 5 <= n0 <= 15                       # 300 moves, one hill per 20 ACCEPTED moves of walker 0
        s2 = sim.engine.get_hill_sigma2(0)
        assert len(s2) == n0 and s2[0] == 2.0 and np.allclose(s2[1:], 2.0 * 1.5 ** 2)
