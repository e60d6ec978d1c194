"""md_pair_grid (two lanes per walker) against md_block_wide as the ensemble grows: where does the pair kernel stop paying?"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from ndsimulator_b200 import config
from ndsimulator_b200.engine import Engine
base = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d", x0=[12.0] * 5,
            integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=False, biases=["mtd"], mtd_w=0.05,
            mtd_sigma=[2.0, 2.0], mtd_dep_freq=100000, mtd_biasf=10.0, mtd_shared=True, precision="f32", steps_per_launch=100,
            mtd_grid=True, mtd_grid_lo=[-8.0, -8.0], mtd_grid_hi=[56.0, 56.0])
for M in (16384, 32768, 49152, 65536):
    rng = np.random.RandomState(1)
    x0 = np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M), rng.uniform(-3, 3, M), np.full(M, 12.0)])
    for label, env in (("pair", None), ("wide", "1")):
        if env: os.environ["NDS_NO_PAIR"] = env
        eng = Engine(config.parse(dict(base, n_walkers=M, steps=100000)).to_struct())
        eng.set_state(x0); eng.begin()
        for _ in range(3): eng.run(100)
        eng.run(2000, sync=True)   # 20 launches of 100 steps, no deposition in between: the step kernel alone
        if env: del os.environ["NDS_NO_PAIR"]
        print(f"M={M} {label}: {eng.last_run_ms() / 20 * 1e3:.1f} us per 100 steps", flush=True)
