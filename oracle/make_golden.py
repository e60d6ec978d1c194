"""TEST INFRASTRUCTURE ONLY: generate ``tests/golden/*.json`` by running the UNMODIFIED Python reference.

Run in the authoring container (where ``/root/reference`` exists):

    python oracle/make_golden.py

The reference is imported through ``oracle/ref_stub.py`` (stand-ins for ``pyfile_utils`` and
``matplotlib`` only -- no arithmetic is replaced).  The fixtures pin the C oracle
(``tests/test_oracle_golden.py``) and, through it, the CUDA path.  NumPy 2.3.5 / Python 3.12.3 were used
for the committed files; floats are written with ``repr`` so they round-trip bit-exactly.
"""
import json
import os
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_stub  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
EXAMPLES = os.path.join(ref_stub.REFERENCE_ROOT, "examples")
TMP = tempfile.mkdtemp()


def tolist(a):
    return np.asarray(a, dtype=float).reshape(-1).tolist()


def dump(name, obj):
    os.makedirs(GOLDEN, exist_ok=True)
    with open(os.path.join(GOLDEN, name), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name)


def unshadow(sim):
    """``AllData.__init__`` (data.py:34) stores ``run.x0`` (the initial position) on every component,
    which shadows the CLASS attribute ``x0`` (well centres) of ``Gaussian2d`` / ``ThreeHole2d/5d``
    (misc.py:12, three_hole.py:17,83): under ``NDRun`` those potentials crash (2-D) or silently use the
    start position as centres (5-D).  The fixtures pin the class-attribute behaviour that the
    reference's own unit test exercises (tests/unit/test_potential.py:19-33, instances built without
    a run), so the shadowing instance attribute is removed here.  Nothing else is touched."""
    pot = sim.potential
    if "x0" in vars(pot) and hasattr(type(pot), "x0"):
        del pot.x0
    return sim


def base_cfg(**kw):
    cfg = dict(root=TMP, run_name="g", dump=False, screen=False, mass=1.0, temperature=300, dt=1.0,
               method="md", seed=0, steps=10)
    cfg.update(kw)
    return cfg


# --------------------------------------------------------------------------- point evaluations
def eval_points():
    rng = np.random.RandomState(2024)
    out = []
    combos = [
        # potential, ndim, colvar, true_colvar, box for sample points
        ("Mueller2d", 2, "Original", "Original", (5.0, 45.0)),
        ("Mueller2d", 2, "ExtractX", "Original", (5.0, 45.0)),
        ("Mueller2d", 2, "ExtractY", "Original", (5.0, 45.0)),
        ("Mueller2d", 2, "XmY", "Original", (5.0, 45.0)),
        ("Mueller2d", 2, "XpY", "Original", (5.0, 45.0)),
        ("Mueller2d", 2, "Rotate2d", "Original", (5.0, 45.0)),
        ("Mueller2d", 2, "Proj2dto1d", "Original", (5.0, 45.0)),
        ("DoubleWell", 2, "Original", "Original", (0.0, 40.0)),
        ("DoubleWell1d", 1, "Original", "Original", (0.0, 40.0)),
        ("Gaussian2d", 2, "Original", "Original", (5.0, 35.0)),
        ("Gaussian", 2, "Original", "Original", (-5.0, 5.0)),
        ("Flat2d", 2, "Original", "Original", (0.0, 10.0)),
        ("Flat1d", 1, "Original", "Original", (0.0, 10.0)),
        ("ThreeHole2d", 2, "Original", "Original", (2.0, 13.0)),
        ("ThreeHole2d", 2, "Rotate2d", "Rotate2d", (2.0, 13.0)),
        ("Mueller5d", 5, "Proj5dto2d", "Proj5dto2d", (5.0, 30.0)),
        ("Mueller5d", 5, "Original", "Proj5dto2d", (5.0, 30.0)),
        ("Mueller5d", 5, "Proj5dto1d", "Proj5dto2dv2", (5.0, 30.0)),
        ("Mueller5d", 5, "Proj5dto1dx0", "Proj5dto2d", (5.0, 30.0)),
        ("Mueller5d", 5, "Proj5dto1dx2", "Proj5dto2d", (5.0, 30.0)),
        ("Mueller5d", 5, "Proj5dto2dv2", "Proj5dto2dv2", (5.0, 30.0)),
        ("Mueller5dshl", 5, "Proj5dto2d", "Proj5dto2d", (5.0, 30.0)),
        ("ThreeHole5d", 5, "Proj5dto2d", "Proj5dto2d", (1.0, 9.0)),
    ]
    for pot, ndim, cv, tcv, (lo, hi) in combos:
        cfg = base_cfg(ndim=ndim, potential=pot, colvar_name=cv, true_colvar_name=tcv,
                       x0=[lo + 1.0] * ndim, integrate="nve")
        sim = unshadow(ref_stub.make_run(cfg))
        pts = rng.uniform(lo, hi, size=(6, ndim))
        rec = dict(config={k: cfg[k] for k in ("ndim", "potential", "colvar_name", "true_colvar_name")},
                   points=[], V=[], F=[], colv=[], J=[])
        for x in pts:
            V, f = sim.potential.compute(x=np.array(x))
            c = np.asarray(sim.colvar.compute(np.array(x)), dtype=float).reshape(-1)
            J = np.asarray(sim.colvar.jacobian(np.array(x)), dtype=float).reshape(len(c), ndim)
            rec["points"].append(tolist(x))
            rec["V"].append(float(V))
            rec["F"].append(tolist(f))
            rec["colv"].append(tolist(c))
            rec["J"].append(tolist(J))
        out.append(rec)
    # the survey's spot values are regenerated here as well
    dump("eval_points.json", out)


def bias_points():
    rng = np.random.RandomState(77)
    out = []
    # umbrella: 2-D Original, 5-D Proj5dto2d, 1-output colvar, two restraints
    umb_cases = [
        dict(ndim=2, potential="Mueller2d", biases=["umb"], umb_k=0.1, umb_r0=[18, 18], x0=[18.0, 18.0]),
        dict(ndim=2, potential="Mueller2d", biases=["umb"], umb_k=[0.1, 0.3], umb_r0=[20, 25],
             colvar_name="Rotate2d", x0=[18.0, 18.0]),
        dict(ndim=2, potential="Mueller2d", biases=["umb"], umb_k=0.2, umb_r0=[3.0],
             colvar_name="XmY", x0=[18.0, 18.0]),
        dict(ndim=5, potential="Mueller5d", biases=["umb"], umb_k=0.1, umb_r0=[18, 18],
             colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d", x0=[18.0, 0.0, 18.0, 0.0, 100.0]),
        dict(ndim=2, potential="Mueller2d", biases=["umb"], umb_n=2, umb_0_k=0.1, umb_0_r0=[18, 18],
             umb_1_k=0.05, umb_1_r0=[25, 10], x0=[18.0, 18.0]),
    ]
    for case in umb_cases:
        cfg = base_cfg(integrate="nve", **case)
        if isinstance(cfg.get("umb_k"), list):  # Harmonic needs an ndarray for per-CV k (umb.py:29)
            cfg["umb_k"] = np.array(cfg["umb_k"])
        sim = ref_stub.make_run(cfg)
        ndim = cfg["ndim"]
        pts = rng.uniform(8.0, 30.0, size=(5, ndim))
        rec = dict(config={k: v for k, v in case.items()}, points=[], Vb=[], Fb=[])
        for x in pts:
            col = sim.colvar.compute(np.array(x))
            Vb, Fb = 0.0, np.zeros(ndim)
            for fix in sim.fixes:
                v, f = fix.compute(x0=np.array(x), col0=col)
                Vb += v
                Fb = Fb + f
            rec["points"].append(tolist(x))
            rec["Vb"].append(float(Vb))
            rec["Fb"].append(tolist(Fb))
        out.append(rec)
    # metadynamics with a given hill table (bias/gaus.py:236-261)
    mtd_cases = [
        dict(ndim=2, potential="Mueller2d", biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 2.0],
             mtd_dep_freq=100, mtd_biasf=10.0, x0=[22.0, 30.0]),
        dict(ndim=2, potential="Mueller2d", biases=["mtd"], mtd_w=0.05, mtd_sigma=[1.5],
             mtd_dep_freq=100, mtd_biasf=10.0, colvar_name="XpY", x0=[22.0, 30.0]),
        dict(ndim=5, potential="Mueller5d", biases=["mtd"], mtd_w=0.05, mtd_sigma=[2.0, 3.0],
             mtd_dep_freq=100, mtd_biasf=10.0, colvar_name="Proj5dto2d",
             true_colvar_name="Proj5dto2d", x0=[12.0] * 5),
        dict(ndim=2, potential="Mueller2d", biases=["mtd", "umb"], mtd_w=0.05, mtd_sigma=[2.0, 2.0],
             mtd_dep_freq=100, mtd_biasf=10.0, umb_k=0.1, umb_r0=[20, 20], x0=[22.0, 30.0]),
        dict(ndim=2, potential="Mueller2d", biases=["pe"], pe_thred=-0.3, x0=[22.0, 30.0]),
    ]
    for case in mtd_cases:
        cfg = base_cfg(integrate="nve", **case)
        sim = ref_stub.make_run(cfg)
        ndim = cfg["ndim"]
        cd = sim.colvar.colvardim
        nh = 37
        centers = rng.uniform(10.0, 30.0, size=(nh, cd))
        heights = rng.uniform(0.01, 0.05, size=nh)
        for fix in sim.fixes:
            if type(fix).__name__ == "Gaussian":
                fix._history = centers.copy()
                fix._w = heights.copy()
        pts = rng.uniform(10.0, 30.0, size=(5, ndim))
        rec = dict(config={k: v for k, v in case.items()}, hills_centers=centers.tolist(),
                   hills_heights=heights.tolist(), points=[], Vb=[], Fb=[])
        for x in pts:
            col = sim.colvar.compute(np.array(x))
            Vb, Fb = 0.0, np.zeros(ndim)
            for fix in sim.fixes:
                v, f = fix.compute(x0=np.array(x), col0=col)
                Vb += v
                Fb = Fb + np.asarray(f).reshape(-1)
            rec["points"].append(tolist(x))
            rec["Vb"].append(float(Vb))
            rec["Fb"].append(tolist(Fb))
        out.append(rec)
    dump("bias_points.json", out)


# ------------------------------------------------------------------------------- trajectories
def run_traj(cfg, sample_every, v0=None):
    sim = unshadow(ref_stub.make_run(cfg, v0=v0))
    sim.begin()
    a = sim.atoms
    rec = dict(config={k: v for k, v in cfg.items() if k not in ("root", "run_name")},
               v0=tolist(a.velocities), begin=dict(pe=float(a.pe), ke=float(a.ke), T=float(a.T),
                                                   totale=float(a.totale)),
               sample_every=sample_every, steps=[], x=[], v=[], f=[], fixf=[], colv=[], pe=[], ke=[],
               biase=[], totale=[], T=[])
    steps = cfg["steps"]
    sim.steps = 0
    done = 0
    while done < steps:
        n = min(sample_every, steps - done)
        sim.steps = done + n
        sim.run()
        done = sim.step
        rec["steps"].append(int(sim.step))
        rec["x"].append(tolist(a.positions))
        rec["v"].append(tolist(a.velocities))
        rec["f"].append(tolist(a.forces))
        rec["fixf"].append(tolist(getattr(a, "fixf", np.zeros_like(a.positions))))  # Minimize never sets fixf
        rec["colv"].append(tolist(a.colv))
        for k in ("pe", "ke", "biase", "totale", "T"):
            rec[k].append(float(getattr(a, k)))
        if sim.method == "committor" and sim.commit_basin != -1:
            break
    rec["final_time"] = float(sim.time)
    for fix in sim.fixes:
        if type(fix).__name__ == "Gaussian" and fix._history is not None:
            rec["hills_centers"] = np.asarray(fix._history, dtype=float).tolist()
            rec["hills_heights"] = tolist(fix._w)
            if fix.adapsig:
                rec["hills_sigma2"] = [float(np.asarray(m).reshape(-1)[0]) for m in fix._sigma_history]
            rec["VR"] = float(fix.VR)
    if sim.method == "committor":
        rec["commit_basin"] = int(sim.commit_basin)
        rec["final_step"] = int(sim.step)
    return rec


def trajectories():
    out = {}
    md2d = yaml.safe_load(open(os.path.join(EXAMPLES, "2d-md.yaml")))
    md2d.update(root=TMP, run_name="g", dump=False, screen=False)
    out["2d_nve_seed7"] = run_traj(dict(md2d, integrate="nve", seed=7, steps=10000), 500)
    out["2d_langevin_gamma0_seed7"] = run_traj(dict(md2d, integrate="langevin", gamma=0.0, seed=7,
                                                    steps=2000), 500)
    out["2d_langevin_gamma0.1_seed0"] = run_traj(dict(md2d, gamma=0.1, seed=0, steps=20000), 1000)
    out["2d_md_yaml_as_shipped_seed0"] = run_traj(dict(md2d, seed=0, steps=2000), 200)
    out["2d_langevin2_seed3"] = run_traj(dict(md2d, integrate="2nd-langevin", gamma=0.01, seed=3,
                                              steps=3000), 300)
    out["2d_rescale_seed3"] = run_traj(dict(md2d, integrate="rescale", rescale_freq=3, seed=3,
                                            x0=[12.0, 12.0], steps=300), 30)
    umb5 = yaml.safe_load(open(os.path.join(EXAMPLES, "5d-umb.yaml")))
    umb5.update(root=TMP, run_name="g", dump=False, screen=False, seed=0, steps=5000)
    out["5d_umb_seed0"] = run_traj(umb5, 500)
    umb2 = yaml.safe_load(open(os.path.join(EXAMPLES, "2d-umb.yaml")))
    umb2.update(root=TMP, run_name="g", dump=False, screen=False, seed=1, steps=3000)
    out["2d_umb_seed1"] = run_traj(umb2, 300)
    mtd = yaml.safe_load(open(os.path.join(EXAMPLES, "2d-mtd.yaml")))
    mtd.update(root=TMP, run_name="g", dump=False, screen=False, seed=0, steps=3000)
    out["2d_mtd_seed0"] = run_traj(mtd, 300)
    out["5d_mtd_seed2"] = run_traj(
        base_cfg(ndim=5, potential="Mueller5d", colvar_name="Proj5dto2d",
                 true_colvar_name="Proj5dto2d", integrate="langevin", gamma=0.1, biases=["mtd"],
                 mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0,
                 x0=[12.0] * 5, seed=2, steps=1500), 150)
    out["1d_doublewell_nve_seed4"] = run_traj(
        base_cfg(ndim=1, potential="DoubleWell1d", x0=30.0, integrate="nve", seed=4, steps=2000,
                 plot_boundary=[[0.0, 40.0]]), 200)
    out["1d_doublewell_langevin_seed4"] = run_traj(
        base_cfg(ndim=1, potential="DoubleWell1d", potential_K1=1.5, x0=30.0, gamma=0.1, seed=4,
                 steps=2000, plot_boundary=[[0.0, 40.0]]), 200)
    out["2d_threehole_nve_seed5"] = run_traj(
        base_cfg(ndim=2, potential="ThreeHole2d", x0=[7.0, 9.0], integrate="nve", seed=5, dt=0.5,
                 temperature=3000, steps=2000), 200)
    out["2d_doublewell_pe_seed6"] = run_traj(
        base_cfg(ndim=2, potential="DoubleWell", x0=[20.0, 12.0], gamma=0.05, seed=6, biases=["pe"],
                 pe_thred=-0.45, steps=1500), 150)
    # method: minimize (engines/minimize.py:41-82): steepest descent with dt * 10.  Minimize.update returns None,
    # so NDRun.run (ndrun.py:220,230) leaves its loop after ONE update per call: sampled after every call
    out["2d_minimize"] = run_traj(dict(md2d, method="minimize", x0=[20.0, 20.0], seed=1, steps=60), 1)
    # adapsig_geo with 1-D colvars (the only case in which gaus.py:263-288 is finite): constant Jacobian
    # (XmY), position-dependent Jacobian (Proj5dto1d, note its 2.0*1e-7 term)
    adap = dict(integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05, mtd_sigma=2.0, mtd_dep_freq=50,
                mtd_biasf=10.0, mtd_adapsig=True, mtd_adapsig_geo=True)
    out["2d_mtd_adapsig_geo_XmY_seed2"] = run_traj(
        base_cfg(ndim=2, potential="Mueller2d", colvar_name="XmY", x0=[22.0, 24.0], seed=2, steps=1200, **adap), 100)
    out["5d_mtd_adapsig_geo_Proj5dto1d_seed3"] = run_traj(
        base_cfg(ndim=5, potential="Mueller5d", colvar_name="Proj5dto1d", true_colvar_name="Proj5dto2d",
                 x0=[12.0] * 5, seed=3, steps=1000, **adap), 100)
    # adapsig WITHOUT adapsig_geo (gaus.py:166-183, 193-212): the variance of every hill is the exponentially weighted
    # variance of the last tao_d = min(50 dep_freq, 100) Stat samples of the colvar.  In the reference this only runs
    # for a colvar of shape (1,) -- ndim 1 with the Original colvar: the (1, 1)-shaped 1-output colvars of two.py /
    # five.py (A28) make dcol.T.dot(exp) fail at the second deposition, and for 2-D colvars the element-wise
    # 1 / covariance of gaus.py:268-277 makes the bias diverge.  Case 1 crosses n > tao_d (a sample every step,
    # tao_d = 100); case 2 samples every 5 steps with a shorter window (dep_freq 1 -> tao_d = 50, a hill every step
    # of time... every dt).
    hist = dict(ndim=1, potential="DoubleWell1d", x0=30.0, integrate="langevin", gamma=0.1, biases=["mtd"],
                mtd_sigma=2.0, mtd_biasf=10.0, mtd_adapsig=True, mtd_adapsig_geo=False, plot_boundary=[[0.0, 40.0]])
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # gaus.py:203 prints the window on every deposition
        out["1d_mtd_adapsig_hist_doublewell_seed5"] = run_traj(
            base_cfg(seed=5, steps=1200, mtd_w=0.05, mtd_dep_freq=50, **hist), 100)
        out["1d_mtd_adapsig_hist_doublewell_stat5_seed6"] = run_traj(
            base_cfg(seed=6, steps=1500, stat_freq=5, mtd_w=0.002, mtd_dep_freq=7, **hist), 100)
    dump("trajectories.json", out)


def committor():
    cm = yaml.safe_load(open(os.path.join(EXAMPLES, "2d-committor.yaml")))
    cm.update(root=TMP, run_name="g", dump=False, screen=False)
    out = dict(config={k: v for k, v in cm.items() if k not in ("root", "run_name")}, shots=[])
    for seed in range(24):
        sim = ref_stub.make_run(dict(cm, seed=seed))
        sim.begin()
        v0 = tolist(sim.atoms.velocities)
        sim.run()
        out["shots"].append(dict(seed=seed, v0=v0, basin=int(sim.commit_basin), step=int(sim.step),
                                 x=tolist(sim.atoms.positions)))
        print(seed, sim.commit_basin, sim.step)
    # diffusion_time (committor.py:100-101): started INSIDE basin 0, a shot may only commit from NDRun.step 37 on
    dcfg = dict(cm, x0=[22.5, 32.5], engine_diffusion_time=37, steps=2000)
    out["diffusion"] = dict(config={k: v for k, v in dcfg.items() if k not in ("root", "run_name")}, shots=[])
    for seed in range(8):
        sim = ref_stub.make_run(dict(dcfg, seed=seed))
        sim.begin()
        v0 = tolist(sim.atoms.velocities)
        sim.run()
        out["diffusion"]["shots"].append(dict(seed=seed, v0=v0, basin=int(sim.commit_basin), step=int(sim.step),
                                              x=tolist(sim.atoms.positions)))
    dump("committor.json", out)
    # temperature 0: the inner engine is Minimize (engines/__init__.py:22-23) -- deterministic relaxation
    t0 = dict(config={k: v for k, v in dict(cm, temperature=0, seed=0, steps=3000).items()
                      if k not in ("root", "run_name")}, runs=[])
    rng = np.random.RandomState(3)
    starts = [[21.0, 17.0], [30.0, 12.0], [24.0, 30.0], [35.0, 8.0], [26.0, 20.0], [22.0, 33.0], [44.0, 5.0]]
    starts += rng.uniform([12.0, 2.0], [46.0, 36.0], size=(17, 2)).tolist()
    for x0 in starts:
        sim = ref_stub.make_run(dict(cm, temperature=0, seed=0, steps=3000, x0=x0))
        sim.begin()
        sim.run()
        t0["runs"].append(dict(x0=tolist(x0), basin=int(sim.commit_basin), step=int(sim.step),
                               x=tolist(sim.atoms.positions), pe=float(sim.atoms.pe),
                               colv=tolist(sim.atoms.colv), time=float(sim.time)))
    dump("committor_T0.json", t0)


def constants():
    """integrator constants (md.py:81-99) and units for the BASELINE configs."""
    out = {}
    for name, kw in {
        "C2_langevin_gamma0.1": dict(ndim=2, potential="Mueller2d", x0=[22.0, 24.0], gamma=0.1),
        "C4_5d_umb": dict(yaml.safe_load(open(os.path.join(EXAMPLES, "5d-umb.yaml")))),
    }.items():
        cfg = base_cfg(**kw)
        cfg.update(root=TMP, run_name="g", dump=False, screen=False)
        sim = ref_stub.make_run(cfg)
        e = sim.engine
        out[name] = dict(m=float(e.m), kBT=float(e.kBT),
                         **{k: float(getattr(e, k)) for k in ("c1", "c2", "c3", "c4", "c5") if hasattr(e, k)})
    dump("constants.json", out)


def path_colvars():
    """Path / Path5dto2d (colvars/path.py): values and Jacobians at random points, plus the reference's
    own known answers (tests/unit/test_colvar_path.py:18,43,65) and a short umbrella trajectory in path
    space.  The reference-point table of tests/inputs/path_ref is stored inside the fixture."""
    ref_stub.install()
    from ndsimulator.colvars import Path, Path5dto2d
    out = {}
    data = np.loadtxt(os.path.join(ref_stub.REFERENCE_ROOT, "tests", "inputs", "path_ref"))
    tables = {
        "kat_1d": dict(ref=[[1], [3], [5], [7]], ref_ind=[0, 1, 2, 3], sigma=1, point=[6.0],
                       expect=0.8287934528751985),
        "kat_2d": dict(ref=[[1, 3.14], [2.05, 4], [3.9, 5], [4.6, 5.9]], ref_ind=[0, 1, 2, 3], sigma=1.0,
                       point=[2.5, 3.1], expect=0.24646052616257966),
        "kat_file": dict(ref=data[:, 1:].tolist(), ref_ind=data[:, 0].tolist(), sigma=1.0,
                         point=[17.40447553, 19.68877675], expect=0.50117643),
    }
    rng = np.random.RandomState(5)
    for name, t in tables.items():
        cov = Path(ref=np.array(t["ref"], dtype=float), ref_ind=np.array(t["ref_ind"], dtype=float).reshape(-1, 1),
                   sigma=t["sigma"])
        cov.initialize()
        nd = len(t["point"])
        pts = [np.array(t["point"], dtype=float)] + [np.array(t["point"]) + rng.normal(scale=1.5, size=nd) for _ in range(6)]
        rec = dict(t, ndim=nd, points=[], value=[], jac=[])
        for x in pts:
            rec["points"].append(tolist(x))
            rec["value"].append(float(cov.compute(x)[0]))
            rec["jac"].append(tolist(cov.jacobian(x)))
        out[name] = rec
    cov5 = Path5dto2d(ref=data[:, 1:], ref_ind=data[:, 0].reshape(-1, 1), sigma=1.5)
    cov5.initialize()
    rec = dict(ref=data[:, 1:].tolist(), ref_ind=data[:, 0].tolist(), sigma=1.5, ndim=5, points=[], value=[], jac=[])
    for _ in range(7):
        x = np.array([rng.uniform(5, 25), rng.uniform(-5, 5), rng.uniform(10, 30), rng.uniform(-5, 5), rng.uniform(0, 50)])
        rec["points"].append(tolist(x))
        rec["value"].append(float(cov5.compute(x)[0]))
        rec["jac"].append(tolist(cov5.jacobian(x)))
    out["path5dto2d_file"] = rec
    # umbrella sampling along the 2-D path variable, NVE (deterministic), through NDRun
    cfg = base_cfg(ndim=2, potential="Mueller2d", x0=[22.0, 28.0], integrate="nve", seed=11, steps=400,
                   colvar_name="Path", colvar_ref=data[:, 1:], colvar_ref_ind=data[:, 0].reshape(-1, 1),
                   colvar_sigma=2.0, biases=["umb"], umb_k=5.0, umb_r0=[0.5])
    t = run_traj(cfg, 50)
    t["config"]["colvar_ref"] = data[:, 1:].tolist()
    t["config"]["colvar_ref_ind"] = data[:, 0].tolist()
    out["umb_on_path_nve"] = t
    dump("path_colvar.json", out)


def dump_files():
    """Text dumps of the reference (io/dump.py) for deterministic runs: NVE and gamma=0 metadynamics."""
    out = {}
    cases = {
        "2d_nve_dump10": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                              integrate="nve", temperature=300, dt=1.0, seed=7, steps=40, dump=True,
                              dump_freq=10, screen=False),
        "2d_mtd_gamma0_dump100": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 30.0],
                                      method="md", integrate="langevin", gamma=0.0, temperature=300,
                                      dt=1.0, seed=3, steps=300, biases=["mtd"], mtd_w=0.05,
                                      mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, dump=True,
                                      dump_freq=100, screen=False),
        "5d_umb_nve_stat10": dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
                                  true_colvar_name="Proj5dto2d", x0=[18.0, 0.0, 18.0, 0.0, 100.0],
                                  biases=["umb"], umb_k=0.1, umb_r0=[18, 18], integrate="nve", dt=1.0,
                                  temperature=300.0, seed=2, steps=60, dump=True, dump_freq=20,
                                  stat_freq=10, screen=False),
        "2d_xmy_nve": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                           integrate="nve", colvar_name="XmY", temperature=300, dt=1.0, seed=7, steps=20,
                           dump=True, dump_freq=10, screen=False),
        # examples/2d-md.yaml as shipped (Langevin), shortened: the stochastic default user path
        "2d_md_yaml_langevin_dump10": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                                           integration="langevin", temperature=300, dt=1.0, seed=0, steps=60,
                                           dump=True, dump_freq=10, screen=False),
        # adaptive variances are listed in fix0.dat instead of sigma (gaus.py:455-456); "[c]" colvar tokens (A28)
        "2d_mtd_adapsig_geo_xmy_gamma0": dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 30.0],
                                               method="md", integrate="langevin", gamma=0.0, temperature=300,
                                               dt=1.0, seed=3, steps=150, colvar_name="XmY", biases=["mtd"],
                                               mtd_w=0.05, mtd_sigma=2.0, mtd_dep_freq=50, mtd_biasf=10.0,
                                               mtd_adapsig=True, mtd_adapsig_geo=True, dump=True, dump_freq=50,
                                               screen=False),
    }
    # umbrella + metadynamics: Gaussian.update recomputes every fix after a deposition (gaus.py:138-146), so the
    # fix_VR column of the umbrella changes at the deposition steps only
    cases["2d_umb_mtd_gamma0_dump20"] = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 30.0], method="md",
                                             integrate="langevin", gamma=0.0, temperature=300, dt=1.0, seed=3,
                                             steps=160, biases=["umb", "mtd"], umb_k=0.02, umb_r0=[20, 28], mtd_w=0.05,
                                             mtd_sigma=[2.0, 2.0], mtd_dep_freq=50, mtd_biasf=10.0, dump=True,
                                             dump_freq=20, screen=False)
    for name, cfg in cases.items():
        d = tempfile.mkdtemp()
        sim = ref_stub.make_run(dict(cfg, root=d, run_name="r"))
        sim.begin()
        v0 = tolist(sim.atoms.velocities)
        sim.run()
        sim.end()
        import logging
        logging.shutdown()
        files = {}
        for fn in sorted(os.listdir(os.path.join(d, "r"))):
            if fn.endswith(".dat"):
                with open(os.path.join(d, "r", fn)) as f:
                    files[fn] = f.read()
        out[name] = dict(config=cfg, v0=v0, files=files)
    dump("dump_files.json", out)


# ------------------------------------------------------------------------------- Metropolis MC
def mc_runs():
    """Metropolis MC runs of the reference made reproducible by seeding the GLOBAL numpy RNG the engine proposes
    with (mc.py:52-55).  The test rebuilds the same uniform numbers with numpy (RandomState(global_seed)
    for randint / rand, RandomState(seed) for the acceptance draws after the ndim velocity draws of begin()), feeds
    them to the oracle and the GPU as caller-supplied uniforms, and compares every recorded move."""
    out = {}
    base = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="mc", dt=0.5, mc_bounds=[1.5, 1.5],
                mc_global_bounds=[48.0, 40.0], temperature=600, dump=False, screen=False, root=TMP, run_name="g",
                steps=400, seed=4, plot_boundary=[[0.0, 48.0], [0.0, 40.0]])
    cases = {
        "single_2d": dict(base),
        "all_2d_umb": dict(base, mc_propose_dim="all", biases=["umb"], umb_k=0.05, umb_r0=[25, 20]),
        "single_2d_mtd": dict(base, biases=["mtd"], mtd_w=0.02, mtd_sigma=[1.5, 1.5], mtd_dep_freq=10, mtd_biasf=8.0),
        "single_2d_mtd_adapsig_geo": dict(base, colvar_name="XmY", biases=["mtd"], mtd_w=0.02, mtd_sigma=1.5,
                                          mtd_dep_freq=10, mtd_biasf=8.0, mtd_adapsig=True, mtd_adapsig_geo=True),
    }
    # Wang-Landau (wang_landau.py:65-161, one basin): same random-number consumption as MC
    cases["wl_single_2d"] = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 30.0], method="wl-mc",
                                 bounds=[2.0, 2.0], global_bounds=[48.0, 40.0], temperature=300, dump=False,
                                 screen=False, root=TMP, run_name="g", steps=400, seed=4,
                                 plot_boundary=[[0.0, 48.0], [0.0, 40.0]])
    # metadynamics under Wang-Landau: fix.update deposits on the accepted-move clock, the acceptance compares
    # pe + biase bins (wang_landau.py:98-126 with fix.energy, gaus.py:296-363)
    cases["wl_single_2d_mtd"] = dict(cases["wl_single_2d"], biases=["mtd"], mtd_w=0.02, mtd_sigma=[1.5, 1.5],
                                     mtd_dep_freq=10, mtd_biasf=8.0, dt=0.5, seed=6)
    for name, cfg in cases.items():
        global_seed = 1000 + len(out)
        sim = unshadow(ref_stub.make_run(cfg))
        np.random.seed(global_seed)
        sim.begin()
        a = sim.atoms
        rec = dict(config={k: v for k, v in cfg.items() if k not in ("root", "run_name")}, global_seed=global_seed,
                   x=[], pe=[], biase=[], accepted=[], time=[])
        for k in range(cfg["steps"]):
            sim.steps = k + 1
            sim.run()
            rec["x"].append(tolist(a.positions))
            rec["pe"].append(float(a.pe))
            rec["biase"].append(float(a.biase))
            rec["accepted"].append(int(sim.engine.accept_rate))
            rec["time"].append(float(sim.time))
        for fix in sim.fixes:
            if type(fix).__name__ == "Gaussian" and fix._history is not None:
                rec["hills_centers"] = np.asarray(fix._history, dtype=float).tolist()
                rec["hills_heights"] = tolist(fix._w)
                rec["VR"] = float(fix.VR)
                if fix.adapsig:
                    rec["hills_sigma2"] = [float(np.asarray(m).reshape(-1)[0]) for m in fix._sigma_history]
        if cfg["method"] == "wl-mc":
            rec["hE"] = tolist(sim.engine.hE[0])
        out[name] = rec
    dump("mc.json", out)


def fes_projection():
    """Gaussian.projection (gaus.py:365-438) on the plot grid np.meshgrid(arange(b0, b1, inc), ...) of io/plot.py:483-488,
    called the way Plot does: once while the run is under way and again at the end, so that the incremental
    bookkeeping (_dep_projection) is the reference's own.  The 1-D adaptive branch (gaus.py:425-436) has no
    `if value is False` guard and re-adds EVERY hill on every call: that case records one call only."""
    out = {}
    mtd = yaml.safe_load(open(os.path.join(EXAMPLES, "2d-mtd.yaml")))
    mtd.update(root=TMP, run_name="g", dump=False, screen=False, seed=0, steps=1000)
    cases = {
        "2d_fixed_sigma_incremental": (mtd, [[0.0, 48.0], [0.0, 40.0]], [1.0, 1.0], 2),
        "1d_fixed_sigma_incremental": (
            base_cfg(ndim=2, potential="Mueller2d", colvar_name="XmY", x0=[22.0, 24.0], seed=2, steps=1000,
                     integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05, mtd_sigma=[1.5], mtd_dep_freq=50,
                     mtd_biasf=10.0), [[-30.0, 30.0]], [0.25], 2),
        "1d_adapsig_geo_single_call": (
            base_cfg(ndim=5, potential="Mueller5d", colvar_name="Proj5dto1d", true_colvar_name="Proj5dto2d",
                     x0=[12.0] * 5, seed=3, steps=600, integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=0.05,
                     mtd_sigma=2.0, mtd_dep_freq=50, mtd_biasf=10.0, mtd_adapsig=True, mtd_adapsig_geo=True),
            [[0.0, 30.0]], [0.2], 1),
    }
    for name, (cfg, boundary, inc, calls) in cases.items():
        sim = unshadow(ref_stub.make_run(cfg))
        sim.begin()
        fix = [f for f in sim.fixes if type(f).__name__ == "Gaussian"][0]
        b = np.array(boundary, dtype=float)
        if len(boundary) == 2:
            X, Y = np.meshgrid(np.arange(b[0, 0], b[0, 1], inc[0]), np.arange(b[1, 0], b[1, 1], inc[1]))
        else:
            X, Y = np.arange(b[0, 0], b[0, 1], inc[0]), None
        total, done, maps, n_at_call = cfg["steps"], 0, [], []
        for c in range(calls):
            sim.steps = total * (c + 1) // calls
            sim.run()
            m = fix.projection(X, Y)
            maps.append(np.array(m, dtype=float).copy())
            n_at_call.append(int(fix._history.shape[0]))
        rec = dict(config={k: v for k, v in cfg.items() if k not in ("root", "run_name")}, boundary=boundary,
                   increment=inc, shape=list(maps[-1].shape), hills_at_call=n_at_call,
                   hills_centers=np.asarray(fix._history, dtype=float).tolist(), hills_heights=tolist(fix._w),
                   maps=[tolist(m) for m in maps])
        if fix.adapsig:
            rec["hills_sigma2"] = [float(np.asarray(m).reshape(-1)[0]) for m in fix._sigma_history]
        out[name] = rec
    dump("fes_projection.json", out)


def map50to2d():
    """Map50to2d (colvars/misc.py:19-156): values and Jacobians of the reference at 24 configurations (unit and
    10-unit scale, plus the coordinate axes that carry the off-diagonal couplings)."""
    ref_stub.install()
    from ndsimulator.colvars.misc import Map50to2d
    cov = Map50to2d()
    cov.ndim = 50  # Colvar.initialize would copy it from the run (colvar.py:11-13); no run can exist
    rng = np.random.RandomState(50)
    pts = [rng.normal(size=50) for _ in range(8)] + [10.0 * rng.normal(size=50) for _ in range(8)]
    for d in (0, 1, 2, 3, 4, 5, 8, 49):
        e = np.zeros(50)
        e[d] = 1.5
        e[(d + 1) % 50] = -0.5
        pts.append(e)
    rec = dict(points=[], value=[], jac=[])
    for x in pts:
        rec["points"].append(tolist(x))
        rec["value"].append(tolist(cov.compute(x)))
        rec["jac"].append(tolist(cov.jacobian(x)))
    dump("map50to2d.json", rec)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("dump_files", "path_colvars", "committor", "mc_runs", "trajectories",
                                             "fes_projection", "map50to2d"):
        globals()[sys.argv[1]]()
        sys.exit(0)
    eval_points()
    bias_points()
    constants()
    trajectories()
    committor()
    dump_files()
    path_colvars()
    mc_runs()
    fes_projection()
    map50to2d()
