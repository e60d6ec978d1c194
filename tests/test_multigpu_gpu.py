"""GPU, world_size 2, NCCL, one process per GPU: the REAL exchange paths of shared-hill metadynamics -- CUDA-IPC peer
stores + the device-side flag barrier, the native ncclAllGather / ncclAllReduce, and both grid-increment exchanges --
checked, not only timed: two ranks with M walkers each must end with the hill table and the walker states of ONE
engine with 2M walkers (the Philox stream is keyed by the global walker id, so the noise is the same).
Needs two GPUs (``gpurun --gpus 2``); skipped on a one-GPU box, where the two-engine emulations of
tests/test_parity_gpu.py and tests/test_mtd_grid_gpu.py cover the host logic."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, NSTEPS = 96, 400
CFG = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d", true_colvar_name="Proj5dto2d",
           x0=[12.0] * 5, integrate="langevin", gamma=0.1, dt=1.0, temperature=300, biases=["mtd"], mtd_w=0.05,
           mtd_sigma=[2.0, 2.0], mtd_dep_freq=25, mtd_biasf=10.0, steps=NSTEPS, mtd_shared=True, seed=77, dump=False,
           steps_per_launch=25)
GRID = dict(mtd_grid=True, mtd_grid_lo=[-8.0, -8.0], mtd_grid_hi=[48.0, 48.0])


def _x0():
    rng = np.random.RandomState(6)
    return np.stack([rng.uniform(8, 30, 2 * M), rng.uniform(-3, 3, 2 * M), rng.uniform(8, 30, 2 * M),
                     rng.uniform(-3, 3, 2 * M), np.full(2 * M, 12.0)])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from ndsimulator_b200 import config, distributed as D
    from ndsimulator_b200.engine import Engine
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        x0 = _x0()
        for precision in ("f64", "f32"):
            for grid in (False, True):
                for mode in ("p2p", "nccl"):
                    cfg = D.shard_config(dict(CFG, precision=precision, **(GRID if grid else {})), n_walkers_per_rank=M)
                    eng = Engine(config.parse(cfg).to_struct())
                    (D.attach_peer_hill_exchange if mode == "p2p" else D.attach_native_hill_exchange)(eng)
                    eng.set_state(x0[:, rank * M:(rank + 1) * M])
                    eng.begin()
                    eng.run(NSTEPS)
                    st = eng.get_state(fields=("x", "v", "thermo"))
                    c, h = eng.get_hills()
                    out = dict(x=st["x"], v=st["v"], biase=st["biase"], hills_c=c, hills_h=h,
                               checksum=np.uint64(eng.table_checksum(0, eng.n_hills())), kernel=eng.kernel_name)
                    if grid:
                        out["grid"] = eng.get_mtd_grid()
                    np.savez(os.path.join(out_dir, f"{precision}_{'grid' if grid else 'exact'}_{mode}_rank{rank}.npz"), **out)
                    dist.barrier()
                    del eng
    finally:
        dist.destroy_process_group()


def test_two_gpus_two_processes_equal_one_engine(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from ndsimulator_b200.engine import Engine
    from tests.helpers import scaled_err, spec_from
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    x0 = _x0()
    exact_runs = {}
    for precision in ("f64", "f32"):
        for grid in (False, True):
            one = Engine(spec_from(dict(CFG, precision=precision, **(GRID if grid else {})), n_walkers=2 * M).to_struct())
            one.set_state(x0)
            one.begin()
            one.run(NSTEPS)
            ref = one.get_state(fields=("x", "v", "thermo"))
            rc, rh = one.get_hills()
            for mode in ("p2p", "nccl"):
                tag = f"{precision}_{'grid' if grid else 'exact'}_{mode}"
                r = [np.load(os.path.join(str(tmp_path), f"{tag}_rank{k}.npz")) for k in range(2)]
                x = np.concatenate([r[0]["x"], r[1]["x"]], axis=1)
                v = np.concatenate([r[0]["v"], r[1]["v"]], axis=1)
                if not grid:
                    # every rank holds the whole table: identical on both ranks, and the single engine's table
                    assert r[0]["checksum"] == r[1]["checksum"], tag
                    assert np.array_equal(r[0]["hills_c"], r[1]["hills_c"]) and np.array_equal(r[0]["hills_h"], r[1]["hills_h"]), tag
                    assert r[0]["hills_h"].shape == rh.shape == ((NSTEPS // 25) * 2 * M,), tag
                    if str(r[0]["kernel"]) == one.kernel_name:  # same kernel variant -> same order of every sum: bit for bit
                        assert np.array_equal(r[0]["hills_c"], rc) and np.array_equal(r[0]["hills_h"], rh), tag
                        assert np.array_equal(x, ref["x"]) and np.array_equal(v, ref["v"]), tag
                    tol = 1e-9 if precision == "f64" else 2e-3
                    assert scaled_err(r[0]["hills_c"], rc) < tol and scaled_err(r[0]["hills_h"], rh) < tol, tag
                    assert scaled_err(x, ref["x"]) < tol, tag
                    exact_runs[(precision, mode)] = (x, r[0]["hills_h"])
                else:
                    # the grid holds every rank's hills (the table only the rank's own); both ranks bit-identical;
                    # against the single engine the increments were summed in another order: rounding level
                    assert np.array_equal(r[0]["grid"], r[1]["grid"]), tag
                    g1 = one.get_mtd_grid()
                    tol = 1e-10 if precision == "f64" else 2e-5
                    assert scaled_err(r[0]["grid"][..., 0], g1[..., 0]) < tol, tag
                    own = np.concatenate([r[0]["hills_h"], r[1]["hills_h"]])
                    assert own.shape == rh.shape and np.allclose(np.sort(own), np.sort(rh), rtol=1e-3 if precision == "f32" else 1e-8), tag
                    assert scaled_err(x, ref["x"]) < (1e-8 if precision == "f64" else 2e-3), tag
    for precision in ("f64", "f32"):  # the two transports move the same records: identical results
        assert np.array_equal(exact_runs[(precision, "p2p")][0], exact_runs[(precision, "nccl")][0]), precision
        assert np.array_equal(exact_runs[(precision, "p2p")][1], exact_runs[(precision, "nccl")][1]), precision
