"""Command line: ``python -m ndsimulator_b200.scripts.run <config.yaml>`` -- the reference's
``ndsimulator <yaml>`` (ndsimulator/scripts/run.py:20-76) on the B200 engine.  Same YAML keys, same
defaults (verbose=info, run_name=ndrun, seed=0, root=./); extra keys: n_walkers, precision,
steps_per_launch, device, dump_walkers, mtd_shared, umb_per_walker.

Under ``torchrun`` (one rank per GPU) ``n_walkers`` is the count PER RANK (the ensemble has world x n_walkers walkers);
give ``n_walkers_global`` instead to have the total divided over the ranks."""
import argparse
import logging
import time

import yaml

from ndsimulator_b200.ndrun import NDRun

default_config = dict(verbose="info", run_name="ndrun", seed=0, root="./")  # scripts/run.py:17


def parse_command_line(args=None):
    parser = argparse.ArgumentParser(description="Run a ND simulation on the B200 multi-walker engine")
    parser.add_argument("config", help="configuration file")
    parser.add_argument("--set", nargs="*", default=[], metavar="KEY=VALUE",
                        help="override YAML keys, e.g. --set n_walkers=65536 precision=f32 (under torchrun n_walkers "
                             "is per rank; n_walkers_global is divided over the ranks)")
    ns = parser.parse_args(args=args)
    config = dict(default_config)
    with open(ns.config) as f:
        config.update(yaml.safe_load(f))
    for kv in ns.set:
        k, v = kv.split("=", 1)
        config[k] = yaml.safe_load(v)
    # one rank per GPU under torchrun: walkers shard by global id (ndsimulator_b200/distributed.py)
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from ndsimulator_b200.distributed import shard_config
        world = int(os.environ["WORLD_SIZE"])
        total = config.pop("n_walkers_global", None)
        if total is not None:
            if int(total) % world:
                raise ValueError(f"n_walkers_global = {total} is not a multiple of the {world} ranks")
            config["n_walkers"] = int(total) // world
        config = shard_config(config)
        config["run_name"] = f"{config['run_name']}_rank{os.environ.get('RANK', '0')}"
    return config


def main(args=None):
    start = time.time()
    config = parse_command_line(args)
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    simulation = NDRun(**config)
    if world > 1 and "mtd" in simulation.biases and simulation.spec.mtd_shared:
        from ndsimulator_b200 import distributed as D
        from ndsimulator_b200.engine import NdsError
        # peer-memory stores need fixed-width records and an IPC mapping; adaptive tables (mtd_adapsig*) and boxes
        # without peer access take the allgather that NCCL provides
        if simulation.spec.mtd_adapsig_geo or simulation.spec.mtd_adapsig_hist:
            D.attach_native_hill_exchange(simulation.engine)
        else:
            try:
                D.attach_peer_hill_exchange(simulation.engine)
            except (NdsError, RuntimeError) as e:
                logging.info(f"peer-memory hill exchange unavailable ({e}); using the NCCL allgather")
                D.attach_native_hill_exchange(simulation.engine)
    simulation.begin()
    simulation.run()
    simulation.end()
    logging.info(f"total time: {time.time() - start}")  # scripts/run.py:63-64
    return simulation


if __name__ == "__main__":
    main()
