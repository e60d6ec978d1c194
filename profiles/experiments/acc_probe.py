import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from ndsimulator_b200 import config
from ndsimulator_b200.engine import Engine
hi = float(sys.argv[1]); M = int(sys.argv[2])
base = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d", x0=[12.0] * 5,
            integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=False, biases=["mtd"], mtd_w=0.05,
            mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, mtd_shared=True, precision="f32", steps_per_launch=100,
            mtd_grid=True, n_walkers=M, steps=2000, mtd_grid_lo=[-8.0, -8.0], mtd_grid_hi=[hi, hi])
rng = np.random.RandomState(1)
x0 = np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M), rng.uniform(-3, 3, M), np.full(M, 12.0)])
eng = Engine(config.parse(base).to_struct()); eng.set_state(x0); eng.begin(); eng.run(600)
