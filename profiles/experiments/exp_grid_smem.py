"""mtd_grid step kernel: grid copy in shared memory (1-D TMA per CTA) vs grid read through L2, 8192 and 2^17 walkers."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from ndsimulator_b200 import config
from ndsimulator_b200.engine import Engine
base = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d", x0=[12.0] * 5,
            integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=False, biases=["mtd"], mtd_w=0.05,
            mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, mtd_shared=True, precision="f32", steps_per_launch=100,
            mtd_grid=True)
for M in (8192, 1 << 17):
    rng = np.random.RandomState(1)
    x0 = np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M), rng.uniform(-3, 3, M), np.full(M, 12.0)])
    for label, box, env in (("113^2 smem", ([-8.0, -8.0], [48.0, 48.0]), None), ("113^2 L2", ([-8.0, -8.0], [48.0, 48.0]), "1"),
                            ("129^2 L2", ([-8.0, -8.0], [56.0, 56.0]), None)):
        if env: os.environ["NDS_GRID_NO_SMEM"] = env
        spec = config.parse(dict(base, n_walkers=M, steps=100 * 60, mtd_grid_lo=box[0], mtd_grid_hi=box[1]))
        eng = Engine(spec.to_struct())
        if env: del os.environ["NDS_GRID_NO_SMEM"]
        eng.set_state(x0); eng.begin()
        for _ in range(5): eng.run(100)
        ms = []
        for _ in range(20):
            eng.run(100, sync=True); ms.append(eng.last_run_ms())
        eng.run(2000, sync=True)
        print(f"M={M} {label}: {np.median(ms)*1e3:.1f} us per stride, back-to-back {eng.last_run_ms()/20*1e3:.1f} us per stride "
              f"-> {M*100/(eng.last_run_ms()/20*1e-3):.3g} walker-steps/s  [{eng.kernel_name}]", flush=True)
