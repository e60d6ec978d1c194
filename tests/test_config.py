"""CPU: YAML / kwargs vocabulary -> nds_config (prefix rules, name resolution, the reference's errors)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import yaml

from ndsimulator_b200 import _abi, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def base(**kw):
    d = dict(ndim=2, potential="Mueller2d", x0=[22.0, 24.0], mass=1.0)
    d.update(kw)
    return d


def test_defaults_follow_reference():
    s = config.parse(dict(ndim=2, x0=[0.0, 0.0]))
    assert s.potential == "Gaussian"            # potentials/_build.py:10
    assert s.colvar == s.true_colvar == "Original"
    assert (s.temperature, s.mass, s.steps, s.method) == (300.0, 5, 100, "md")  # ndrun.py:80-86
    assert (s.dt, s.gamma, s.integrate) == (0.5, 0.005, "langevin")            # md.py:50-52
    assert s.initial_temperature == 300.0


def test_prefix_rules():
    s = config.parse(base(gamma=0.5, md_gamma=0.01))
    assert s.gamma == 0.01
    s = config.parse(base(gamma=0.5, md_gamma=0.01, engine_gamma=0.02))
    assert s.gamma == 0.02                      # later prefixes win
    s = config.parse(base(integration="nve"))   # not a constructor parameter: ignored
    assert s.integrate == "langevin"
    s = config.parse(base(biases=["umb"], umb_n=2, umb_k=0.1, umb_r0=[1, 2], umb_1_r0=[3, 4]))
    assert len(s.umb_r0) == 2 and s.umb_k[1].tolist() == [0.1, 0.1]
    s = config.parse(base(biases=["mtd"], mtd_w=0.05, mtd_sigma=2.0, mtd_dep_freq=100, mtd_biasf=10.0))
    assert s.mtd_sigma.tolist() == [2.0, 2.0]   # float sigma broadcast (gaus.py:76-78)
    assert s.mtd_capacity >= 2
    s = config.parse(base(dump_freq=10, stat_freq=5))
    assert (s.dump_freq, s.stat_freq) == (10, 5)


def test_example_yaml_keys():
    p = "/root/reference/examples"
    if not os.path.isdir(p):
        pytest.skip("reference examples not present")
    for name in ("2d-md.yaml", "2d-committor.yaml", "2d-mtd.yaml", "2d-umb.yaml", "5d-umb.yaml", "1d-md.yaml"):
        cfg = yaml.safe_load(open(os.path.join(p, name)))
        s = config.parse(cfg)
        st = s.to_struct()
        assert st.ndim == cfg["ndim"]
    s = config.parse(yaml.safe_load(open(os.path.join(p, "5d-umb.yaml"))))
    assert s.integrate == "2nd-langevin" and s.gamma == 0.01 and s.colvar == "Proj5dto2d"
    s = config.parse(yaml.safe_load(open(os.path.join(p, "2d-committor.yaml"))))
    assert s.method == "committor" and s.n_basins == 2 and s.criteria.shape == (2, 2, 2)


def test_reference_errors():
    with pytest.raises(NameError):
        config.parse(base(method="bogus"))                       # engines/__init__.py:28
    with pytest.raises(NameError):
        config.parse(base(x0=[1.0, 2.0, 3.0]))                   # ndrun.py:168
    with pytest.raises(AssertionError):
        config.parse(base(ndim=5, x0=[0.0] * 5))                 # potential.py:12
    with pytest.raises(AssertionError):
        config.parse(base(integrate="verlet"))                   # md.py:59
    with pytest.raises(NameError):
        config.parse(base(biases=["mtd"], mtd_w=0.1, mtd_sigma=[1.0], mtd_biasf=5.0))  # gaus.py:82
    with pytest.raises(AssertionError):
        config.parse(base(biases=["umb"], umb_k=0.1, umb_r0=[1.0]))                    # umb.py:18
    with pytest.raises(TypeError):
        config.parse(base(integrate="rescale"))                  # md.py:220
    with pytest.raises(ValueError):
        config.parse(base(potential="mueller2d"))                # potentials/_build.py:18 (stale yaml, A14)
    with pytest.raises(ValueError):
        config.parse(base(biases=["mtd"], mtd_w=0.1, mtd_sigma=[1.0, 1.0]))  # biasf=1 divides by zero (A3)
    with pytest.raises(AssertionError):
        config.parse(base(method="committor", n_basins=2, criteria=[[[0, 1], [0, 1]]]))  # committor.py:57


def test_out_of_scope_is_refused_loudly():
    with pytest.raises(NotImplementedError):
        config.parse(base(method="read_dump"))
    assert config.parse(base(method="wl-mc", bounds=[1, 1])).to_struct().method == _abi.NDS_METHOD_WL
    assert config.parse(base(method="mc", mc_bounds=[1, 1])).method == "mc"
    # temperature 0 selects the Minimize inner engine (engines/__init__.py:22-23)
    s = config.parse(base(method="committor", temperature=0, n_basins=1, criteria=[[[0, 1], [0, 1]]]))
    assert s.integrate == "minimize" and s.emin == -1.2 and s.to_struct().min_maxiter == 1000
    assert config.parse(base(method="minimize")).integrate == "minimize"
    with pytest.raises(ValueError):  # history-based adapsig with a 2-D colvar: divergent in the reference (gaus.py:268-277)
        config.parse(base(biases=["mtd"], mtd_w=0.1, mtd_sigma=[1.0, 1.0], mtd_biasf=2.0, mtd_adapsig=True))
    h = config.parse(base(colvar_name="XmY", biases=["mtd"], mtd_w=0.1, mtd_sigma=1.5, mtd_biasf=2.0,
                          mtd_adapsig=True)).to_struct()
    assert h.mtd_adapsig_hist == 1 and h.mtd_adapsig_geo == 0
    with pytest.raises(ValueError):  # adapsig_geo with a 2-D colvar: inf / NaN in the reference (gaus.py:268-275)
        config.parse(base(biases=["mtd"], mtd_w=0.1, mtd_sigma=[1.0, 1.0], mtd_biasf=2.0, mtd_adapsig=True,
                          mtd_adapsig_geo=True))
    a = config.parse(base(colvar_name="XmY", biases=["mtd"], mtd_w=0.1, mtd_sigma=1.5, mtd_biasf=2.0,
                          mtd_adapsig=True, mtd_adapsig_geo=True)).to_struct()
    assert a.mtd_adapsig_geo == 1 and a.mtd_sigma[0] == 1.5


def test_user_defined_python_plugins_are_rejected():
    for key in ("colvar_name", "true_colvar_name", "potential"):
        with pytest.raises(ValueError, match="cannot run on the B200 engine"):
            config.parse(base(**{key: "my_module.MyClass"}))
    with pytest.raises(ValueError, match="cannot run on the B200 engine"):
        config.parse(base(potential=lambda x: x))
    assert config.parse(base(potential="ndsimulator.potentials.Mueller2d")).potential == "Mueller2d"


def test_struct_matches_header():
    src = '#include <stdio.h>\n#include "include/ndsb200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(nds_config), ' \
          '__builtin_offsetof(nds_config, criteria), __builtin_offsetof(nds_config, stat_freq));return 0;}'
    exe = os.path.join("/tmp", f"nds_sz_{os.getpid()}")
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["gcc", "-x", "c", "-", "-I", ROOT, "-o", exe], input=src.encode(), cwd=ROOT, env=env,
                   check=True)
    size, off_crit, off_stat = map(int, subprocess.check_output([exe]).split())
    os.remove(exe)
    assert C.sizeof(_abi.NdsConfig) == size
    assert _abi.NdsConfig.criteria.offset == off_crit
    assert _abi.NdsConfig.stat_freq.offset == off_stat


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports everything include/ndsb200.h declares; compute
    calls are not made here.  nds_create must fail loudly (no CPU fallback) when there is no device."""
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _abi.load_library()
    import re
    hdr = open(os.path.join(ROOT, "include", "ndsb200.h")).read()
    declared = set(re.findall(r"NDS_API [^;]*?\b(nds_[a-z_0-9]+)\(", hdr))
    assert declared and declared == set(_abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    import torch
    if not torch.cuda.is_available():
        st = config.parse(base()).to_struct()
        h = C.c_void_p()
        assert lib.nds_create(C.byref(st), C.byref(h)) != 0
        assert b"no CPU fallback" in lib.nds_last_error(None)
        from ndsimulator_b200 import Engine, NdsError
        with pytest.raises(NdsError):
            Engine(st)


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under ndsimulator_b200/ (Python or CUDA / C++ sources) may import,
    link or mention it, and the engine wrapper must fail loudly when the CUDA library is missing."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ndsimulator_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(d, fn), errors="ignore").read()
                if re.search(r"\boracle\b", text):
                    offenders.append(os.path.join(d, fn))
    assert not offenders, offenders
    src = open(os.path.join(pkg, "_abi.py")).read()
    assert "raise" in src and "libndsb200.so" in src  # load() raises when the library is missing: no fallback
