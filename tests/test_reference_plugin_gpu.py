"""GPU: the UNMODIFIED reference ``NDRun`` driving the B200 engine through the engine plugin (``method: b200``).

``ndsimulator_b200/plugin.py`` implements the reference's engine protocol (``engines/__init__.py:11-31``,
``ndrun.py:152,186,230,236-240``); ``plugin.register()`` is the three-line factory patch of INTEGRATION.md applied
at run time.  The reference's own loop, ``Stat`` and ``Dump`` then write the files -- which must be the files the
reference writes with ``method: md`` (tests/golden/dump_files.json).  Needs an importable reference: the
authoring container's /root/reference or the copy ``oracle/fetch_ref.py`` installs under baseline/_ref (that one
travels to the GPU box)."""
import os

import numpy as np
import pytest

from tests.helpers import ROOT, load_golden
from tests.test_frontend_gpu import assert_same_text

pytestmark = pytest.mark.gpu


def _reference_root():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "ndsimulator")):
            return p
    return None


@pytest.fixture(scope="module")
def ref_stub():
    root = _reference_root()
    if root is None:
        pytest.skip("no importable reference (baseline/_ref is installed by __graft_entry__.build())")
    from oracle import ref_stub as rs
    rs.REFERENCE_ROOT = root
    rs.install()
    from ndsimulator_b200 import plugin
    plugin.register()
    return rs


@pytest.mark.parametrize("name", ["2d_nve_dump10", "2d_mtd_gamma0_dump100", "5d_umb_nve_stat10", "2d_xmy_nve",
                                  "2d_mtd_adapsig_geo_xmy_gamma0", "2d_umb_mtd_gamma0_dump20"])
@pytest.mark.parametrize("block", [7, 256])
def test_reference_ndrun_with_b200_engine_writes_the_reference_files(ref_stub, name, block, tmp_path):
    rec = load_golden("dump_files.json")[name]
    cfg = dict(rec["config"], method="b200", root=str(tmp_path), run_name="r", screen=False, steps_per_launch=block,
               n_walkers=5, mtd_shared=False)   # every walker its own hills: walker 0 is the reference's run
    sim = ref_stub.make_run(cfg, v0=rec["v0"])
    from ndsimulator_b200.plugin import B200Engine
    assert isinstance(sim.engine, B200Engine) and type(sim).__module__ == "ndsimulator.ndrun"
    sim.begin()
    sim.run()
    sim.end()
    import logging
    for fn, want in rec["files"].items():
        path = os.path.join(str(tmp_path), "r", fn)
        for h in logging.getLogger(path).handlers:
            h.flush()
        with open(path) as f:
            got = f.read()
        assert_same_text(got, want, what=f"{name}/{fn}")
    st = sim.engine.get_state()
    assert st["x"].shape == (cfg["ndim"], 5) and np.all(np.isfinite(st["x"]))
    assert np.allclose(st["x"][:, 0], sim.atoms.positions, rtol=1e-12)  # walker 0 is run.atoms
