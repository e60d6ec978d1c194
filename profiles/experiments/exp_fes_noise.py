"""How many independent runs does the 0.1 kT FES bar need?  GPU (fp64 / fp32) vs oracle ensembles, rms and SEM in kT."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.getcwd())
from ndsimulator_b200.engine import Engine
from ndsimulator_b200.constant import kB
from oracle import oracle
from tests.helpers import spec_from

kT = kB * 300
MD2D = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], dt=1.0, temperature=300)

def single(R, nsteps, w):
    cfg = dict(MD2D, x0=[22.0, 30.0], integrate="langevin", gamma=0.1, biases=["mtd"], mtd_w=w,
               mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, steps=nsteps, mtd_shared=False)
    spec = spec_from(cfg, n_walkers=R, precision="f64")
    st = spec.to_struct()
    x0 = np.repeat(spec.x0.reshape(-1, 1), R, axis=1)
    t0 = time.time()
    eng = Engine(st); eng.set_state(x0); eng.begin(); eng.run(nsteps)
    ga_all = np.array([eng.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40, walker=w_) for w_ in range(R)])
    t1 = time.time()
    orc = oracle.OracleSim(st, seeds=np.arange(R)); orc.set_threads(oracle.max_threads()); orc.set_state(x0); orc.begin(); orc.run(nsteps)
    gb_all = np.array([orc.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40, walker=w_) for w_ in range(R)])
    t2 = time.time()
    ga, gb = ga_all.mean(0), gb_all.mean(0)
    mask = gb > 2 * kT
    d = (ga - gb)[mask]; d -= d.mean()
    sem = gb_all[:, mask].std(axis=0).mean() / np.sqrt(R)
    print(f"single R={R} n={nsteps} w={w}: mask {mask.sum()} rms {np.sqrt((d**2).mean())/kT:.4f} kT, sem {sem/kT:.4f} kT, gpu {t1-t0:.1f}s oracle {t2-t1:.1f}s", flush=True)

def c5(M, n, nseeds, w=0.002):
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d", x0=[12.0] * 5,
               integrate="langevin", gamma=0.1, dt=1.0, temperature=300, biases=["mtd"], mtd_w=w, mtd_sigma=[2.0, 2.0],
               mtd_dep_freq=100, mtd_biasf=10.0, steps=n, mtd_shared=True)
    rng = np.random.RandomState(0)
    x0 = np.stack([rng.uniform(14, 26, M), rng.uniform(-2, 2, M), rng.uniform(20, 32, M), rng.uniform(-2, 2, M), np.full(M, 12.0)])
    t0 = time.time()
    G = []
    for s in range(nseeds):
        eng = Engine(spec_from(dict(cfg, seed=1 + s), n_walkers=M, precision="f32", steps_per_launch=100).to_struct())
        eng.set_state(x0); eng.begin(); eng.run(n)
        G.append(eng.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40))
    t1 = time.time()
    O = []
    for s in range(nseeds):
        orc = oracle.OracleSim(spec_from(cfg, n_walkers=M).to_struct(), seeds=np.arange(1000 * (s + 1), 1000 * (s + 1) + M))
        orc.set_threads(oracle.max_threads()); orc.set_state(x0); orc.begin(); orc.run(n)
        O.append(orc.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40))
    t2 = time.time()
    G, O = np.array(G), np.array(O)
    mask = O.mean(0) > 2 * kT
    d = (G.mean(0) - O.mean(0))[mask]; d -= d.mean()
    semo = O[:, mask].std(axis=0).mean() / np.sqrt(nseeds)
    semg = G[:, mask].std(axis=0).mean() / np.sqrt(nseeds)
    print(f"c5 M={M} n={n} w={w} seeds={nseeds} kernel={eng.kernel_name}: mask {mask.sum()} rms {np.sqrt((d**2).mean())/kT:.4f} kT, sem oracle {semo/kT:.4f} gpu {semg/kT:.4f} kT, gpu {t1-t0:.1f}s oracle {t2-t1:.1f}s", flush=True)


def c5_variants(M, n, nseeds, w):
    cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d", x0=[12.0] * 5,
               integrate="langevin", gamma=0.1, dt=1.0, temperature=300, biases=["mtd"], mtd_w=w, mtd_sigma=[2.0, 2.0],
               mtd_dep_freq=100, mtd_biasf=10.0, steps=n, mtd_shared=True)
    rng = np.random.RandomState(0)
    x0 = np.stack([rng.uniform(14, 26, M), rng.uniform(-2, 2, M), rng.uniform(20, 32, M), rng.uniform(-2, 2, M), np.full(M, 12.0)])
    sets = {}
    for label, prec, env, seed_base in (("f32_mtd_block", "f32", None, 1), ("f32_generic", "f32", "NDS_NO_MTD_BLOCK", 1),
                                        ("f64_mtd_block", "f64", None, 1), ("f64_generic", "f64", "NDS_NO_MTD_BLOCK", 1),
                                        ("f64_generic_otherseeds", "f64", "NDS_NO_MTD_BLOCK", 5001)):
        if env: os.environ[env] = "1"
        G = []
        for s in range(nseeds):
            eng = Engine(spec_from(dict(cfg, seed=seed_base + s), n_walkers=M, precision=prec, steps_per_launch=100).to_struct())
            eng.set_state(x0); eng.begin(); eng.run(n)
            G.append(eng.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40))
        print(label, eng.kernel_name, flush=True)
        if env: del os.environ[env]
        sets[label] = np.array(G)
    for label, base in (("oracle_a", 1000), ("oracle_b", 700000)):
        O = []
        for s in range(nseeds):
            orc = oracle.OracleSim(spec_from(cfg, n_walkers=M).to_struct(), seeds=np.arange(base * (s + 1), base * (s + 1) + M))
            orc.set_threads(oracle.max_threads()); orc.set_state(x0); orc.begin(); orc.run(n)
            O.append(orc.fes_grid(0.0, 1.0, 48, 0.0, 1.0, 40))
        sets[label] = np.array(O)
    mask = sets["oracle_a"].mean(0) > 2 * kT
    names = list(sets)
    for k in names:
        print(f"  {k}: sem {sets[k][:, mask].std(axis=0).mean() / np.sqrt(nseeds) / kT:.4f} kT  mean level {sets[k].mean(0)[mask].mean() / kT:.3f} kT")
    for i, a in enumerate(names):
        for b in names[i + 1:]:
            d = (sets[a].mean(0) - sets[b].mean(0))[mask]
            off = d.mean(); d = d - off
            print(f"  rms {a} vs {b}: {np.sqrt((d**2).mean()) / kT:.4f} kT (offset {off / kT:+.4f})")

print("threads", oracle.max_threads())
c5_variants(64, 3000, 300, 0.008)
