"""The reference's own integration tests (tests/integration/test_md.py, test_mc.py, test_minimize.py,
test_committor.py with the fixtures of tests/conftest.py:13-61), run unchanged in spirit against the drop-in:
same keyword arguments, same begin / run / end protocol, plus the assertions the reference does not make."""
import os
from copy import deepcopy

import numpy as np
import pytest
import yaml

from ndsimulator_b200 import NDRun

pytestmark = pytest.mark.gpu

EX_2D_MD = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md", integration="langevin",
                temperature=300, dt=1.0, steps=100000, dump=True, dump_freq=10, movie=False, oneplot=True,
                plot=True, plot_boundary=[[0.0, 48.0], [0.0, 40.0]], plot_ebound=[-1.4, 0.2], plot_freq=100,
                plot_increment=[1.0, 1.0], verbose="debug", screen=False)  # examples/2d-md.yaml


def basic_md(tmp):  # tests/conftest.py:13-36
    return dict(root=tmp, method="md", ndim=2, potential="Mueller2d", steps=10, colvar_name="Original",
                true_colvar_name="Original", integrate="rescale", rescale_freq=3, dt=1.0, temperature=300,
                mass=1.0, x0=[12.0, 12.0], plot=False, movie=False, dump=True, dump_freq=1,
                boundary=[[2, 40], [2, 40]], verbose="debug")


def basic_mc(tmp):  # tests/conftest.py:39-61
    return dict(basic_md(tmp), method="mc", mc_bounds=[1, 1], mc_global_bounds=[50, 50])


@pytest.mark.parametrize("integrate", ["rescale", "langevin", "2nd-langevin", "nve"])
def test_simple_md(tmp_path, integrate):  # tests/integration/test_md.py:7-21
    config = dict(EX_2D_MD, root=str(tmp_path), run_name=integrate)
    config["steps"] = config["steps"] // 100
    config["integration"] = integrate      # the key the reference test sets; MD ignores it
    run = NDRun(**config)
    run.begin()
    run.run()
    run.end()
    assert run.step == 1000 and run.spec.integrate == "langevin"
    rows = open(os.path.join(str(tmp_path), integrate, "pos.dat")).read().strip().split("\n")
    assert len(rows) == 101 and np.isfinite(np.array(rows[-1].split(), dtype=float)).all()


def test_wholemcprocess_and_propose_all(tmp_path):  # tests/integration/test_mc.py:6-23
    for name, extra in (("mc", {}), ("mc_nve", {"propose_dim": "all"})):
        kwargs = deepcopy(basic_mc(str(tmp_path)))
        kwargs.update(run_name=name, screen=False, **extra)
        run = NDRun(**kwargs)
        run.begin()
        run.run()
        run.end()
        assert run.step == 10 and 0 <= run.accept_rate <= 10


def test_min(tmp_path):  # tests/integration/test_minimize.py:6-14
    kwargs = deepcopy(basic_md(str(tmp_path)))
    kwargs.update(method="minimize", run_name="minimize", screen=False)
    run = NDRun(**kwargs)
    run.begin()
    pe0 = run.atoms.pe
    run.run()
    run.end()
    assert run.atoms.pe < pe0               # steepest descent lowers the energy


def test_simple_committor(tmp_path):  # tests/integration/test_committor.py:5-13
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    config = yaml.safe_load(open(os.path.join(root, "examples", "c3-2d-committor.yaml")))
    config.update(root=str(tmp_path), screen=False, n_walkers=1, precision="f64")
    config["steps"] = config["steps"] // 100
    run = NDRun(**config)
    run.begin()
    run.run()
    run.end()
    assert run.commit_basin in (-1, 0, 1) and run.step <= 100
