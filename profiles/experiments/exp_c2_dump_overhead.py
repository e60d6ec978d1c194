import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from ndsimulator_b200 import config
from ndsimulator_b200.engine import Engine
base = dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md", integrate="langevin", gamma=0.1,
            temperature=300, dt=1.0, seed=0, n_walkers=1<<20, precision="f32", steps_per_launch=1000)
variants = {
  "plain": dict(dump=False),
  "hist": dict(dump=False, dump_freq=10, hist_n=[48,40], hist_lo=[0.,0.], hist_hi=[48.,40.]),
  "frames4096": dict(dump=True, dump_freq=10, stat_freq=10, dump_walkers=4096, dump_capacity=104),
  "both": dict(dump=True, dump_freq=10, stat_freq=10, dump_walkers=4096, dump_capacity=104, hist_n=[48,40], hist_lo=[0.,0.], hist_hi=[48.,40.]),
  "both_nosplit": dict(dump=True, dump_freq=10, stat_freq=10, dump_walkers=4096, dump_capacity=104, hist_n=[48,40], hist_lo=[0.,0.], hist_hi=[48.,40.]),
}
M = 1<<20
for name, extra in variants.items():
    if name == "both_nosplit": os.environ["NDS_NO_SPLIT"] = "1"
    spec = config.parse(dict(base, **extra))
    eng = Engine(spec.to_struct())
    eng.set_state(np.repeat(spec.x0.reshape(-1,1), M, axis=1))
    eng.begin()
    ms = []
    for i in range(8):
        eng.run(1000, sync=True)
        if spec.dump: eng.lib.nds_drain_frames(eng.h, None, 0, None)
        ms.append(eng.last_run_ms())
    print(name, eng.kernel_name, "ms", np.round(ms[3:], 4), "w-s/s %.4g" % (M*1000/(np.mean(ms[3:])*1e-3)))
