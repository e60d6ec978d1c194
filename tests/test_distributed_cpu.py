"""CPU, world_size 2, gloo: the host-side logic of the N > 1 path -- sharding of walkers (global walker
ids), the hill-segment exchange layout of shared metadynamics, and the commit-count reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ndsimulator_b200 import config
from ndsimulator_b200.distributed import exchange_segment, gather_commit_counts, shard_config


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, cd = 6, 2
        cfg = shard_config(dict(ndim=5, potential="Mueller5d", colvar_name="Proj5dto2d",
                                true_colvar_name="Proj5dto2d", x0=[12.0] * 5, mass=1.0, biases=["mtd"],
                                mtd_w=0.05, mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0,
                                steps=1000, dt=1.0, dump=False), n_walkers_per_rank=M)
        spec = config.parse(cfg)
        st = spec.to_struct()
        assert st.walker_offset == rank * M and st.n_walkers_global == world * M and st.device == rank
        assert st.mtd_capacity >= 11 * world * M     # every rank's hills fit the shared table

        # one deposition stride: record = (c0, c1, w, pad) float32 as in the fp32 engine
        stride = 4
        rec = np.zeros((world * M, stride), dtype=np.float32)
        gid = rank * M + np.arange(M)
        rec[gid, 0] = gid + 0.25
        rec[gid, 1] = -gid
        rec[gid, 2] = 0.05 / (1 + gid)
        seg = torch.from_numpy(rec.view(np.uint8).reshape(-1))
        bytes_per_rank = M * stride * 4
        exchange_segment(seg, bytes_per_rank, rank * bytes_per_rank)
        full = seg.numpy().view(np.float32).reshape(world * M, stride)
        g = np.arange(world * M)
        assert np.array_equal(full[:, 0], (g + 0.25).astype(np.float32))
        assert np.array_equal(full[:, 1], (-g).astype(np.float32))
        assert np.allclose(full[:, 2], 0.05 / (1 + g))

        basin = np.array([-1, 0, 1, 1, 0, 1]) if rank == 0 else np.array([1, 1, -1, -1, 0, 0])
        counts = gather_commit_counts(basin, 2)
        assert counts.tolist() == [3, 4, 5]           # uncommitted, basin 0, basin 1 over both ranks
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_shard_config_single_process(monkeypatch):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    c = shard_config(dict(n_walkers=10))
    assert (c["walker_offset"], c["n_walkers_global"], c["device"]) == (0, 10, 0)
