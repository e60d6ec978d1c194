"""CPU: the reference-format dump writer (ndsimulator_b200/dump.py) on synthetic frames -- row layout,
the lagged thermo sample, "[value]" tokens of (1,1)-shaped colvars (A28), the hill listing
of fix{i}.dat (bias/gaus.py:440-461).  The numeric content is checked on the GPU against the files the
reference wrote (tests/test_frontend_gpu.py::test_dump_files_match_reference)."""
import os

import numpy as np

from ndsimulator_b200 import config
from ndsimulator_b200.dump import DumpWriter

NAMES = ["step", "time", "x0", "x1", "v0", "v1", "f0", "f1", "c0", "c1", "T", "pe", "ke", "biase", "totale", "VR"]


def frame(step, nw=2, cd=2):
    names = NAMES if cd == 2 else [n for n in NAMES if n != "c1"]
    fr = np.zeros((len(names), nw))
    for i, n in enumerate(names):
        fr[i] = step * 100 + i + np.arange(nw) * 0.5
    fr[names.index("step")] = step
    fr[names.index("time")] = float(step)
    return fr, names


def read(tmp, rel):
    with open(os.path.join(tmp, rel)) as f:
        return f.read().split("\n")


def test_rows_lag_and_walker_directories(tmp_path):
    spec = config.parse(dict(ndim=2, potential="Mueller2d", x0=[1.0, 2.0], mass=1.0, dump=True, dump_freq=10,
                             stat_freq=1, dump_walkers=2, n_walkers=4, root=str(tmp_path), run_name="r"))
    w = DumpWriter(spec, 2)
    w.open_files()
    frames = []
    for s in (0, 9, 10, 19, 20):
        fr, names = frame(s)
        frames.append(fr)
    w.write_frames(np.array(frames), names)
    w.close_files()
    pos = read(tmp_path, "r/pos.dat")
    assert len([ln for ln in pos if ln]) == 3                       # steps 0, 10, 20 only
    assert pos[1].split() == [repr(1002.0), repr(1003.0)]           # x0, x1 of walker 0 at step 10
    thermo = read(tmp_path, "r/thermo.dat")
    assert thermo[0] == "#time dt T pe ke biase totale"
    assert thermo[1].split()[0] == "0" and thermo[1].split()[5] == "0"   # int 0 time / biase without fixes
    assert thermo[2].split()[0] == "9.0"                            # lagged stat sample (step 9) beside step 10
    assert float(thermo[2].split()[2]) == 900 + names.index("T")
    assert float(thermo[3].split()[0]) == 19.0
    pos1 = read(tmp_path, "r/walker_1/pos.dat")                     # second walker in its own directory
    assert pos1[1].split() == [repr(1002.5), repr(1003.5)]


def test_bracket_colvars_and_hill_listing(tmp_path):
    spec = config.parse(dict(ndim=2, potential="Mueller2d", x0=[1.0, 2.0], mass=1.0, colvar_name="XmY",
                             biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0], mtd_dep_freq=10, mtd_biasf=10.0,
                             dump=True, dump_freq=10, stat_freq=10, dump_walkers=1, root=str(tmp_path),
                             run_name="m", steps=20))
    w = DumpWriter(spec, 1)
    w.open_files()
    frames = []
    for s in (0, 10, 20):
        fr, names = frame(s, nw=1, cd=1)
        frames.append(fr)
    hills = (np.array([[1.5], [2.5]]), np.array([0.05, 0.04]))
    w.write_frames(np.array(frames), names, hills=hills, hills_at=lambda step: {0: 0, 10: 1, 20: 2}[step])
    w.close_files()
    col = read(tmp_path, "m/colvar.dat")
    assert col[0] == str(np.array([8.0])) and col[0].startswith("[")       # "[8.]" like f"{c}" of a (1,) array
    fix = read(tmp_path, "m/fix0.dat")
    assert fix[0] == "" and fix[1] == "0 1.5 2.0 0.05" and fix[2] == "" and fix[3] == "1 2.5 2.0 0.04"
    thermo = read(tmp_path, "m/thermo.dat")
    assert thermo[0].endswith("fix_VR_0") and len(thermo[1].split()) == 8
    assert thermo[1].split()[5] != "0"                                     # biase is a float when fixes exist
