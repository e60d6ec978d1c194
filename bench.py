#!/usr/bin/env python
"""Benchmark of the hot path: walker-steps/s of the fused multi-walker MD kernel (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c5] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: ONE fused launch block of `md_steps_per_step` MD
steps over all walkers of the rank (C2: 2^20 walkers x 1000 MD steps = 1.05e9 walker-steps).  The default
K=100 is the whole C2 job (10^5 MD steps).  value = walker-steps/s with the walker state resident in
HBM; e2e = the same metric through the public API with HOST buffers (pinned H2D of x, v and D2H of
x, v, thermo inside the timed region, every step).

N > 1: launched by torchrun, one rank per GPU; walkers shard with no data-path collective (weak scaling,
C2/C4) or share a hill table through one NCCL allgather per deposition stride (C5).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to its workers.  With that setting libgomp runs the CPU arm's
# `parallel for num_threads(n)` regions at single-core speed (measured here and on the GPU box: 9.8e6 instead of
# 1.1e8 walker-steps/s), which would flatter the GPU/CPU ratio at N > 1.  The CPU baseline must see the same
# OpenMP environment as in the N = 1 run, so the variable is dropped before libgomp is loaded (by numpy / torch).
if os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ.pop("OMP_NUM_THREADS")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_WALKER_STEP = {"c2": 128.0, "c3": 128.0, "c4": 289.0}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` captures (profiles/r01_*_raw.csv) at the default walker counts
NCU_DRAM_BYTES_PER_LAUNCH = {"c2": 22.77e6, "c4": 7.37e6, "c5": 1.59e6}  # SURVEY.md section 8d (algorithmic, fixed)
L2_FLUSH_BYTES = 256 << 20


def workload_cfg(name, n_walkers, precision):
    if name == "c2":  # BASELINE.json configs[1]
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                    integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=False,
                    n_walkers=n_walkers, precision=precision, steps_per_launch=1000)
    if name == "c3":  # examples/2d-committor.yaml: shooting from one seed point, 2^20 shots per GPU
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[21.0, 17.0], method="committor",
                    n_basins=2, criteria=[[[20, 25], [30, 35]], [[38, 50], [2, 10]]], gamma=0.1,
                    temperature=300, dt=1.0, seed=0, dump=False, n_walkers=n_walkers, precision=precision,
                    steps_per_launch=1000, steps=10000)
    if name == "dump":  # C2 physics with EVERY walker dumping a full frame every 10 steps (io/dump.py rows)
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                    integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=True,
                    dump_freq=10, stat_freq=10, dump_walkers=n_walkers, dump_capacity=12,
                    n_walkers=n_walkers, precision=precision, steps_per_launch=100, steps=100)
    if name == "c4":  # examples/5d-umb.yaml, windows x walkers as one batch
        return dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
                    true_colvar_name="Proj5dto2d", x0=[18.0, 0.0, 18.0, 0.0, 100.0], biases=["umb"],
                    umb_k=0.1, umb_r0=[18, 18], umb_per_walker=True, integrate="2nd-langevin",
                    md_gamma=0.01, temperature=300.0, dt=1.0, seed=0, dump=False, n_walkers=n_walkers,
                    precision=precision, steps_per_launch=1000)
    if name == "c5":  # 5-D multiple-walker metadynamics (derived config)
        return dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
                    true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate="langevin", gamma=0.1,
                    temperature=300, dt=1.0, seed=0, dump=False, biases=["mtd"], mtd_w=0.05,
                    mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, mtd_shared=True,
                    n_walkers=n_walkers, precision=precision, steps_per_launch=100)
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [t.strip() for t in s.split(",")]
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
                for n, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(max(smax)) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def initial_state(name, spec, M, rank_offset):
    x0 = np.repeat(spec.x0.reshape(-1, 1), M, axis=1)
    extra = {}
    if name == "c5":
        # SURVEY 8d names x0 = (12,...,12) for every walker; 8192 identical starts would stack 8192 hills
        # (410 eV) on one point at step 0, so the walkers are spread over the Mueller valley instead
        rng = np.random.RandomState(1234 + rank_offset)
        x0 = np.stack([rng.uniform(8, 30, M), rng.uniform(-3, 3, M), rng.uniform(8, 30, M),
                       rng.uniform(-3, 3, M), np.full(M, 12.0)])
    if name == "c4":
        # 16 x 16 windows: r0 on xi1 in {10..40} x xi2 in {6..36}
        gid = rank_offset + np.arange(M)
        win = gid % 256
        r1 = 10.0 + 2.0 * (win % 16)
        r2 = 6.0 + 2.0 * (win // 16)
        x0 = np.stack([r1, np.zeros(M), r2, np.zeros(M), np.full(M, 100.0)])
        extra["k"] = np.full((2, M), 0.1)
        extra["r0"] = np.stack([r1, r2])
    return x0, extra


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would time
    the CPU arm on one core: the oracle gets an explicit thread count instead (num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(name, threads, target_seconds=12.0):
    """The oracle (scalar fp64 C port of the reference path) timed on the host cores of this box."""
    from oracle import oracle
    from ndsimulator_b200 import config
    threads = threads or host_threads()
    M = 256 * threads
    spec = config.parse(workload_cfg(name, M, "f64"))
    x0, extra = initial_state(name, spec, M, 0)

    def run(nsteps):
        o = oracle.OracleSim(spec.to_struct(), seeds=np.arange(M))
        o.set_threads(threads)
        o.set_state(x0)
        if extra:
            o.set_walker_bias(extra["k"], extra["r0"])
        o.begin()
        t = time.perf_counter()
        o.run(nsteps)
        return time.perf_counter() - t

    probe_steps = 200
    dt = run(probe_steps)
    nsteps = int(max(probe_steps, min(2000000, probe_steps * target_seconds / max(dt, 1e-6))))
    dt = run(nsteps)
    return dict(value=M * nsteps / dt, unit="walker-steps/s", cores=threads, kind="port",
                sample=f"{M} walkers x {nsteps} MD steps of workload {name} (fp64 scalar C oracle, "
                       f"OpenMP over walkers, {dt:.1f} s)")


def reference_arm(args):
    """--impl reference: the reference's CPU path (C oracle port; the Python reference itself cannot
    travel to the GPU box) on all host cores, same config / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from ndsimulator_b200 import config
    name = args.workload
    threads = host_threads()
    M = 64 * threads
    md_steps = 1000 if name != "c5" else 100
    spec = config.parse(workload_cfg(name, M, "f64"))
    x0, extra = initial_state(name, spec, M, 0)
    o = oracle.OracleSim(spec.to_struct(), seeds=np.arange(M))
    o.set_threads(threads)
    o.set_state(x0)
    if extra:
        o.set_walker_bias(extra["k"], extra["r0"])
    o.begin()
    for _ in range(args.warmup):
        o.run(md_steps)
    t = time.perf_counter()
    for _ in range(args.steps):
        o.run(md_steps)
    dt = time.perf_counter() - t
    value = M * md_steps * args.steps / dt
    sample = f"{M} walkers x {md_steps} MD steps per step (fp64 scalar C oracle, {threads} OpenMP threads)"
    print(json.dumps({
        "impl": "reference", "metric": "walker-steps/s", "value": value, "unit": "walker-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "walkers": M, "md_steps_per_step": md_steps},
        "cpu_baseline": dict(value=value, unit="walker-steps/s", cores=threads, kind="port", sample=sample),
        "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "dump"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: per workload)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="C5 hill exchange: peer-memory stores fused into the deposit kernel, or NCCL allgather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from ndsimulator_b200 import config
    from ndsimulator_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    name = args.workload
    if name == "c3":
        args.steps = 10 - args.warmup  # the whole job is 10 blocks of 1000 steps (steps: 10000)
    M = args.walkers or {"c2": 1 << 20, "c3": 1 << 20, "c4": 1 << 17, "c5": 8192, "dump": 1 << 20}[name]
    cfg = workload_cfg(name, M, args.precision)
    cfg.update(device=local_rank, walker_offset=rank * M, n_walkers_global=world * M)
    md_steps = cfg["steps_per_launch"]
    if name == "dump":
        args.no_e2e = True
    if name == "c5":
        cfg["steps"] = md_steps * (args.steps + args.warmup + (0 if args.no_e2e else args.steps)) + md_steps
    spec = config.parse(cfg)
    eng = Engine(spec.to_struct())
    if name == "c5" and world > 1:
        from ndsimulator_b200.distributed import attach_hill_exchange, attach_peer_hill_exchange
        if args.exchange == "p2p":
            attach_peer_hill_exchange(eng)
        else:
            attach_hill_exchange(eng)
    x0, extra = initial_state(name, spec, M, rank * M)
    eng.set_state(x0)
    if extra:
        eng.set_walker_bias(extra["k"], extra["r0"])
    eng.begin()

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def drop_frames():
        if name == "dump":  # frames stay on the device; the ring is only rewound (no D2H in the timed region)
            eng.lib.nds_drain_frames(eng.h, None, 0, None)

    for _ in range(args.warmup):
        eng.run(md_steps)
        drop_frames()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)

    # ---- device-resident timing: K steps, CUDA events on the engine's stream, L2 flushed between steps
    barrier()
    launches0 = eng.launch_count
    dev_ms = 0.0
    hills_evals = 0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        if name == "c5":
            hills_evals += M * md_steps * (eng.n_hills() + world * M)  # table after this step's deposit
        eng.run(md_steps, sync=True)
        dev_ms += eng.last_run_ms()
        drop_frames()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    launches = eng.launch_count - launches0
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    value = world * M * md_steps * args.steps / (dev_ms * 1e-3)
    c3_info = None
    if name == "c3":
        # count the steps that were actually integrated in the timed blocks (stopped walkers do nothing)
        basin, exit_step = eng.get_commit()
        lo = args.warmup * md_steps
        done = np.clip(exit_step, lo, None) - lo
        tot = torch.tensor([float(done.sum()), float((basin >= 0).sum()), float((basin == 1).sum())],
                           dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        value = float(tot[0]) / (dev_ms * 1e-3)
        c3_info = {"shots": world * M, "committed": int(tot[1]), "to_basin_1": int(tot[2]),
                   "p_basin1_given_committed": float(tot[2] / max(tot[1], 1)),
                   "integrated_fraction_of_nominal_steps": float(tot[0]) / (world * M * md_steps * args.steps)}
        args.no_e2e = True

    # ---- end to end through the public API with host buffers (pinned), every step
    e2e = None
    if not args.no_e2e:
        nd = spec.ndim
        # two pinned buffer sets, ping-pong: the state read back from step n is the input of step n+1.
        # Host buffers are in the engine's precision (nds_set_state_f32 / nds_get_state_f32 for fp32).
        import ctypes as C
        f32 = args.precision == "f32"
        tdt, cty, esz = (torch.float32, C.c_float, 4) if f32 else (torch.float64, C.c_double, 8)
        bufs = [dict(x=torch.empty((nd, M), dtype=tdt).pin_memory(),
                     v=torch.empty((nd, M), dtype=tdt).pin_memory()) for _ in range(2)]
        st = eng.get_state(fields=("x", "v"))
        bufs[0]["x"].numpy()[:] = st["x"]
        bufs[0]["v"].numpy()[:] = st["v"]
        cp = C.POINTER(cty)
        set_fn = eng.lib.nds_set_state_f32 if f32 else eng.lib.nds_set_state
        get_fn = eng.lib.nds_get_state_f32 if f32 else eng.lib.nds_get_state

        def ptr(tens):
            return C.cast(tens.data_ptr(), cp)

        class Batch:
            """One ensemble with its two pinned host buffer sets (ping-pong: the state read back from pass n
            is the input of pass n+1)."""

            def __init__(self, engine, first):
                self.eng, self.flip = engine, 0
                self.bufs = first if first is not None else [
                    dict(x=torch.empty((nd, M), dtype=tdt).pin_memory(),
                         v=torch.empty((nd, M), dtype=tdt).pin_memory()) for _ in range(2)]
                self.oth = torch.empty((5, M), dtype=tdt).pin_memory()

            def submit(self):
                src = self.bufs[self.flip]
                # positions AND velocities are supplied, so the run continues (no re-draw, no re-begin)
                self.eng._check(set_fn(self.eng.h, ptr(src["x"]), ptr(src["v"])))            # H2D
                self.eng._check(self.eng.lib.nds_run(self.eng.h, md_steps))                  # async launch

            def collect(self):
                dst = self.bufs[1 - self.flip]
                self.eng._check(get_fn(self.eng.h, ptr(dst["x"]), ptr(dst["v"]), None, None, None,
                                       ptr(self.oth)))                                        # D2H (blocks)
                self.flip = 1 - self.flip

        def timed(fn, n):
            barrier()
            t0 = time.perf_counter()
            fn(n)
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])

        batches = [Batch(eng, bufs)]

        def serial(n):
            for _ in range(n):
                batches[0].submit()
                batches[0].collect()

        def pipelined(n):
            # two ensembles in flight: while one integrates, the other's D2H / H2D run on the copy engines
            batches[0].submit()
            for i in range(n):
                if i + 1 < n:
                    batches[(i + 1) % 2].submit()
                batches[i % 2].collect()

        e2e_n = max(4, min(args.steps, 20))
        serial(2)
        serial_s = timed(serial, e2e_n)
        e2e = {"value": world * M * md_steps * e2e_n / serial_s, "unit": "walker-steps/s",
               "h2d_bytes_per_step": 2 * nd * M * esz, "d2h_bytes_per_step": (2 * nd + 5) * M * esz,
               "steps": e2e_n, "mode": "one ensemble, copy -> integrate -> copy in series"}
        if name in ("c2", "c4"):
            # second, independent ensemble (its own walker ids -> its own Philox streams)
            cfg2 = dict(cfg, walker_offset=(world + rank) * M, n_walkers_global=2 * world * M)
            eng2 = Engine(config.parse(cfg2).to_struct())
            x02, extra2 = initial_state(name, spec, M, (world + rank) * M)
            eng2.set_state(x02)
            if extra2:
                eng2.set_walker_bias(extra2["k"], extra2["r0"])
            eng2.begin()
            eng2.run(md_steps)
            b2 = Batch(eng2, None)
            st2 = eng2.get_state(fields=("x", "v"))
            b2.bufs[0]["x"].numpy()[:] = st2["x"]
            b2.bufs[0]["v"].numpy()[:] = st2["v"]
            batches.append(b2)
            pipelined(4)
            pipe_s = timed(pipelined, 2 * e2e_n)
            e2e = dict(e2e, value=world * M * md_steps * 2 * e2e_n / pipe_s, steps=2 * e2e_n,
                       serial_value=e2e["value"],
                       mode="two independent ensembles in flight (double-buffered): every pass still copies its "
                            "inputs H2D from pinned memory and its result D2H; serial_value = one ensemble")

    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks = measured_peaks()
        props = torch.cuda.get_device_properties(local_rank)
        sm_count = props.multi_processor_count
        sm_max_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp32_peak_tflops = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        per_gpu = value / world
        out = {
            "metric": "walker-steps/s", "value": value, "unit": "walker-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": {"c2": "C2: 2-D Mueller-Brown Langevin gamma=0.1, 2^20 walkers/GPU",
                                    "c3": "C3: 2d-committor.yaml shooting, 2^20 shots/GPU, <= 10^4 steps each "
                                          "(a step = the first 1000-step blocks; stopped walkers idle their lane)",
                                    "c4": "C4: 5-D umbrella windows (5d-umb.yaml), 256 windows, 2nd-order Langevin",
                                    "c5": "C5: 5-D multiple-walker metadynamics, shared hill table",
                                    "dump": "C2 physics, every walker records a full frame every 10 steps "
                                            "(device-side dump rows of io/dump.py)"}[name],
                       "walkers_per_gpu": M, "md_steps_per_step": md_steps,
                       "kernel": eng.kernel_name,
                       **({"hill_exchange": args.exchange if world > 1 else "none"} if name == "c5" else {}),
                       "l2": "flushed between steps (256 MiB memset)",
                       "timing": "CUDA events on the engine stream per step, summed; max over ranks",
                       "wall_ms_total": wall_ms},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if name == "dump":
            width = eng.frame_width
            frames_per_step = md_steps // 10
            nbytes = frames_per_step * width * M * 8
            gbs = nbytes / (dev_ms / args.steps * 1e-3) / 1e9
            peak = peaks.get("hbm_gbs", 6650.0)
            out["roofline"] = {
                "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": None,
                "note": f"algorithmic dump bytes: {frames_per_step} frames x {width} fp64 fields x {M} walkers per "
                        f"launch of {md_steps} steps; peak = MEASURED_PEAKS.json hbm_gbs "
                        f"({'measured' if 'hbm_gbs' in peaks else 'fallback'}); the launch also integrates the steps"}
            out["dump_gb_per_s"] = gbs
        elif name in FLOPS_PER_WALKER_STEP:
            ach = per_gpu * FLOPS_PER_WALKER_STEP[name] / 1e12
            out["roofline"] = {
                "bound": "fp32_fma", "achieved": ach, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                "frac": ach / fp32_peak_tflops,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(name) if (args.walkers == 0 and args.precision == "f32") else None,
                "note": f"algorithmic {FLOPS_PER_WALKER_STEP[name]:.0f} flop/walker-step (SURVEY 8d) x "
                        f"walker-steps/s per GPU; peak = {sm_count} SMs x 128 lanes x 2 x {sm_max_mhz:.0f} MHz "
                        "(nvidia-smi clocks.max.sm); MEASURED_PEAKS.json holds no FP32 figure. HBM traffic is "
                        "negligible (state is register-resident for 1000 steps)."}
        else:
            he = world * hills_evals / (dev_ms * 1e-3)
            mufu_peak = sm_count * 16 * sm_max_mhz * 1e6
            out["roofline"] = {
                "bound": "mufu_ex2", "achieved": he / 1e12, "peak": mufu_peak / 1e12 * world,
                "unit": "T hill-evals/s", "frac": he / (mufu_peak * world),
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(name) if args.walkers == 0 else None,
                "note": "one MUFU.EX2 per walker-hill pair; peak = SMs x 16 lanes x clocks.max.sm"}
            out["hill_evals_per_s"] = he
        if c3_info:
            out["committor"] = c3_info
        if e2e:
            out["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(name, 0)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
