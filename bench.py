#!/usr/bin/env python
"""Benchmark of the hot path: walker-steps/s of the fused multi-walker MD kernels (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5|c5_grid|dump]
                    [--impl ours|reference] [--no-workloads]

Headline (default `--workload c2`): a "step" is one pass of the hot path over one batch = ONE fused launch of
`md_steps_per_step` MD steps over all walkers of the rank (C2: 2^20 walkers x 1000 MD steps = 1.05e9 walker-steps);
K = 100 is the whole C2 job (10^5 MD steps).  value = walker-steps/s with the walker state resident in HBM;
e2e = the same metric through the C ABI with pinned HOST buffers, one block per direction, H2D of (x, v) and
D2H of (x, v, thermo) inside the timed region every pass (nds_submit / nds_wait).

The same JSON line carries `workloads`: device-timed records of the other BASELINE configurations -- C3 (committor),
C4 (5-D umbrella windows), C5 (5-D multiple-walker metadynamics, the only path with a collective: exact hill sum with
both exchanges, and the grid-accumulated bias) and the dump path -- each with its own roofline fraction; at N > 1 the
C5 records carry `exchange_check`, an in-run verification that every rank holds the same hills / grid.

N > 1: launched by torchrun, one rank per GPU; walkers shard with no data-path collective (C2/C3/C4), the C5 ranks
share a hill table (peer-memory stores over NVLink or one NCCL allgather per deposition stride) or a bias grid
(peer-memory slots or one NCCL all-reduce per stride).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to its workers.  With that setting libgomp runs the CPU arm's
# `parallel for num_threads(n)` regions at single-core speed, which would flatter the GPU/CPU ratio at N > 1.
if os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ.pop("OMP_NUM_THREADS")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ALGORITHMIC work per walker-step, SURVEY.md section 8d (add / mul = 1, FMA = 2, MUFU = 1)
FLOPS_PER_WALKER_STEP = {"c2": 128.0, "c2_nodump": 128.0, "c3": 128.0, "c4": 289.0, "dump": 128.0, "dump1": 128.0}
# C5 with mtd_grid: colvar 19 + Mueller 92 + chain rule 5 + Langevin 60 + Box-Muller 36 (= the 212 of section 8d)
# + bicubic Hermite interpolation: 2 x 8 basis ops + 2 x 4 x 7 column contractions + 3 x 7 row contractions + 4
GRID_FLOPS_PER_WALKER_STEP = 212.0 + 97.0
# Issue-slot model: issue cycles per 32 walker-steps per SM sub-partition that the executed
# instruction mix needs when an IMAD.WIDE holds the issue port for 4.1 cycles and a packed FP32 instruction for 2.05
# (profiles/microbench/pipes_b200.txt).  The mixes come from the ncu source pages named here (profiles/opmix.py);
# they describe the kernels of this commit and have to be re-measured when a kernel changes.
ISSUE_MODEL = {
    "c2_nodump": (101.6, "profiles/r02_c2_nodump_raw.csv + source page: 58.9 instr, 6.0 IMAD.WIDE, 23.0 packed FP"),
    "c4": (369.6, "profiles/r02_c4_raw.csv + source page: 276.4 instr, 30.1 IMAD.WIDE"),
}
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, from ONE
# `ncu --set full` capture each -- profiles/capture_r02*.sh, raw pages under profiles/ -- at the bench's own sizes.
# Measured numbers of THIS commit's kernels, not derived ones: re-capture when a kernel changes.  For the step
# kernels the algorithmic traffic is the state load at launch entry (2^20 x (x, v) floats = 16.8 MB for C2) plus the
# part of the state written back that reaches DRAM before the kernel ends; the rest stays in L2.
NCU_TRAFFIC = {
    "c2": (16.737e6 + 5.486e6 + 0.095e6, "profiles/r02_c2_raw.csv (hist kernel + frames kernel of one step)"),
    "c2_nodump": (16.800e6 + 6.057e6, "profiles/r02_c2_nodump_raw.csv"),
    "c3": (24.165e6 + 10.055e6, "profiles/r02_c3_raw.csv"),
    "c4": (59.324e6 + 56.141e6, "profiles/r02_c4_raw.csv"),
    "dump1": (16.828e6 + 6714.608e6, "profiles/r02_dump1_raw.csv"),
}
L2_FLUSH_BYTES = 256 << 20
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def workload_cfg(name, n_walkers, precision):
    if name == "c2":  # BASELINE.json configs[1] with the side work SURVEY.md section 8d lists for C2: colvar
        # histogram on the device (all walkers, every 10 steps) + full state of walkers 0..4095 every 10 steps
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                    integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0,
                    dump=True, dump_freq=10, stat_freq=10, dump_walkers=min(4096, n_walkers), dump_capacity=104,
                    hist_n=[48, 40], hist_lo=[0.0, 0.0], hist_hi=[48.0, 40.0],
                    n_walkers=n_walkers, precision=precision, steps_per_launch=1000)
    if name == "c2_nodump":  # the same launch without frames / histogram (round-1 headline)
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                    integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=False,
                    n_walkers=n_walkers, precision=precision, steps_per_launch=1000)
    if name == "c3":  # examples/2d-committor.yaml: shooting from one seed point, 2^20 shots per GPU
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[21.0, 17.0], method="committor",
                    n_basins=2, criteria=[[[20, 25], [30, 35]], [[38, 50], [2, 10]]], gamma=0.1,
                    temperature=300, dt=1.0, seed=0, dump=False, n_walkers=n_walkers, precision=precision,
                    steps_per_launch=1000, steps=10000)
    if name in ("dump", "dump1"):  # C2 physics with EVERY walker dumping a full frame every 10 steps (io/dump.py
        # rows), or every step (dump1: the reference's default dump_freq = 1, io/dump.py:34-39 -- HBM-bound)
        freq = 10 if name == "dump" else 1
        return dict(ndim=2, mass=1.0, potential="Mueller2d", x0=[22.0, 24.0], method="md",
                    integrate="langevin", gamma=0.1, temperature=300, dt=1.0, seed=0, dump=True,
                    dump_freq=freq, stat_freq=freq, dump_walkers=n_walkers, dump_capacity=100 // freq + 2,
                    n_walkers=n_walkers, precision=precision, steps_per_launch=100, steps=100)
    if name == "c4":  # examples/5d-umb.yaml, windows x walkers as one batch
        return dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
                    true_colvar_name="Proj5dto2d", x0=[18.0, 0.0, 18.0, 0.0, 100.0], biases=["umb"],
                    umb_k=0.1, umb_r0=[18, 18], umb_per_walker=True, integrate="2nd-langevin",
                    md_gamma=0.01, temperature=300.0, dt=1.0, seed=0, dump=False, n_walkers=n_walkers,
                    precision=precision, steps_per_launch=1000)
    if name in ("c5", "c5_grid"):  # 5-D multiple-walker metadynamics (derived config)
        cfg = dict(ndim=5, mass=1.0, potential="Mueller5d", colvar_name="Proj5dto2d",
                   true_colvar_name="Proj5dto2d", x0=[12.0] * 5, integrate="langevin", gamma=0.1,
                   temperature=300, dt=1.0, seed=0, dump=False, biases=["mtd"], mtd_w=0.05,
                   mtd_sigma=[2.0, 2.0], mtd_dep_freq=100, mtd_biasf=10.0, mtd_shared=True,
                   n_walkers=n_walkers, precision=precision, steps_per_launch=100)
        if name == "c5_grid":  # bias grid over xi in [-8, 56]^2 at spacing sigma / 4 (129 x 129 nodes, 266 KB)
            cfg.update(mtd_grid=True, mtd_grid_lo=[-8.0, -8.0], mtd_grid_hi=[56.0, 56.0])
        return cfg
    raise SystemExit(f"unknown workload {name}")


WORKLOAD_TEXT = {
    "c2": "C2: 2-D Mueller-Brown Langevin gamma=0.1, 2^20 walkers/GPU",
    "c3": "C3: 2d-committor.yaml shooting, 2^20 shots/GPU, <= 10^4 steps each (a step = the first 1000-step "
          "blocks; stopped walkers idle their lane)",
    "c4": "C4: 5-D umbrella windows (5d-umb.yaml), 256 windows x 4096 walkers, 2nd-order Langevin",
    "c5": "C5: 5-D multiple-walker metadynamics, shared hill table, exact hill sum",
    "c5_grid": "C5: 5-D multiple-walker metadynamics, grid-accumulated bias (mtd_grid)",
    "c2_nodump": "C2 without frames / histogram",
    "dump": "C2 physics, every walker records a full frame every 10 steps (device-side dump rows of io/dump.py)",
    "dump1": "C2 physics, every walker records a full frame EVERY step (the reference's default dump_freq = 1)",
}


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
    if name in ("c5", "c5_grid"):
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
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------ CPU arms
def port_baseline(name, threads, target_seconds=12.0):
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


def reference_cfg(name):
    """The workload as kwargs of the reference's own NDRun (one walker per process)."""
    cfg = workload_cfg(name, 1, "f64")
    for k in ("n_walkers", "precision", "steps_per_launch", "umb_per_walker", "mtd_shared", "mtd_grid",
              "mtd_grid_lo", "mtd_grid_hi", "dump_walkers", "dump_capacity"):
        cfg.pop(k, None)
    # stat_freq 10: the Stat history stays far below its 10^5-sample cap over the whole run
    cfg.update(dump=False, screen=False, plot=False, plot_boundary=[[0.0, 48.0], [0.0, 40.0]], stat_freq=10)
    return cfg


_REF_SIM = None


def _ref_worker_init(name, ref_root, seed_base, tmp):
    """Worker process of the reference arm: ONE unmodified NDRun, kept alive between steps."""
    global _REF_SIM
    os.environ["NDS_REFERENCE_ROOT"] = ref_root
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    import multiprocessing as mp
    from oracle import ref_stub
    ref_stub.REFERENCE_ROOT = ref_root
    ident = mp.current_process()._identity
    idx = ident[0] if ident else 0
    cfg = reference_cfg(name)
    cfg.update(seed=seed_base + idx, steps=0, root=tmp, run_name=f"ref{idx}")
    sim = ref_stub.make_run(cfg)
    sim.begin()
    _REF_SIM = sim


def _ref_worker_step(nsteps):
    sim = _REF_SIM
    sim.steps = sim.step + nsteps   # NDRun.run() advances until self.step == self.steps (ndrun.py:214-221)
    t = time.perf_counter()
    sim.run()
    return nsteps, time.perf_counter() - t


def python_reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "ndsimulator"))


class PythonReference:
    """the UNMODIFIED reference (installed under baseline/_ref by oracle/fetch_ref.py), one
    NDRun per host core, dump / screen / plot off, distinct seeds; reports sum(steps) / wall."""

    def __init__(self, name, cores=None):
        import multiprocessing as mp
        import tempfile
        self.cores = cores or host_threads()
        self.tmp = tempfile.mkdtemp(prefix="nds_ref_")
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_ref_worker_init, initargs=(name, REF_DIR, 1000, self.tmp))

    def step(self, nsteps):
        t = time.perf_counter()
        res = self.pool.map(_ref_worker_step, [nsteps] * self.cores, chunksize=1)
        return sum(r[0] for r in res), time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()
        import shutil
        shutil.rmtree(self.tmp, ignore_errors=True)


def python_reference_baseline(name, target_seconds=10.0):
    ref = PythonReference(name)
    try:
        ref.step(200)               # imports, first-call overheads
        n, dt = ref.step(2000)      # rate estimate
        per = max(2000, int(2000 * target_seconds / max(dt, 1e-3)))
        per = min(per, 60000)       # stay below Stat.mem = 10^5 samples in total
        n, dt = ref.step(per)
    finally:
        ref.close()
    return dict(value=n / dt, unit="walker-steps/s", cores=ref.cores, kind="reference",
                sample=f"{ref.cores} processes x 1 NDRun x {per} MD steps of workload {name}: the unmodified "
                       f"Python reference from baseline/_ref (pyfile_utils / matplotlib stand-ins only), {dt:.1f} s")


def cpu_baseline(name):
    """Reference CPU path beside the GPU number: the unmodified Python reference when it travelled with the snapshot
    (baseline/_ref), with the C port of the same path next to it; the port alone otherwise."""
    port = port_baseline(name, 0)
    if python_reference_available():
        try:
            ref = python_reference_baseline(name)
            ref["port"] = port
            return ref
        except Exception as ex:  # pragma: no cover
            port["reference_error"] = repr(ex)
    return port


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same workload /
    metric.  A step = every core advances its own NDRun by `md_steps` MD steps (a bounded sample of the GPU step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload if args.workload in ("c3", "c4", "c5") else "c2_nodump"
    threads = host_threads()
    if python_reference_available() and not args.port_reference:
        md_steps = 1000
        ref = PythonReference(name, threads)
        try:
            for _ in range(min(args.warmup, 3)):
                ref.step(md_steps)
            t = time.perf_counter()
            total = 0
            for _ in range(args.steps):
                n, _dt = ref.step(md_steps)
                total += n
            dt = time.perf_counter() - t
        finally:
            ref.close()
        value = total / dt
        M = threads
        kind = "reference"
        sample = (f"{threads} processes (one per host core) x 1 unmodified NDRun x {md_steps} MD steps per step, "
                  "Python reference from baseline/_ref")
        extra = {"port": port_baseline(name, threads, target_seconds=6.0)}
    else:
        from oracle import oracle
        from ndsimulator_b200 import config
        M = 64 * threads
        md_steps = 1000 if name != "c5" else 100
        spec = config.parse(workload_cfg(name, M, "f64"))
        x0, extra_state = initial_state(name, spec, M, 0)
        o = oracle.OracleSim(spec.to_struct(), seeds=np.arange(M))
        o.set_threads(threads)
        o.set_state(x0)
        if extra_state:
            o.set_walker_bias(extra_state["k"], extra_state["r0"])
        o.begin()
        for _ in range(args.warmup):
            o.run(md_steps)
        t = time.perf_counter()
        for _ in range(args.steps):
            o.run(md_steps)
        dt = time.perf_counter() - t
        value = M * md_steps * args.steps / dt
        kind = "port"
        sample = f"{M} walkers x {md_steps} MD steps per step (fp64 scalar C oracle, {threads} OpenMP threads)"
        extra = {}
    print(json.dumps({
        "impl": "reference", "metric": "walker-steps/s", "value": value, "unit": "walker-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[name], "walkers": M, "md_steps_per_step": md_steps},
        "cpu_baseline": dict(value=value, unit="walker-steps/s", cores=threads, kind=kind, sample=sample, **extra),
        "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """Process-wide plumbing of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.peaks = measured_peaks()
        props = torch.cuda.get_device_properties(self.local_rank)
        self.sm_count = props.multi_processor_count
        self.sm_max_mhz = self.peaks.get("sm_max_mhz") or 1965.0

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(v) for v in t]

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def fp32_peak_tflops(self):
        return self.sm_count * 128 * 2 * self.sm_max_mhz * 1e6 / 1e12


def make_engine(ctx, name, M, precision, **over):
    from ndsimulator_b200 import config
    from ndsimulator_b200.engine import Engine
    cfg = workload_cfg(name, M, precision)
    cfg.update(device=ctx.local_rank, walker_offset=ctx.rank * M, n_walkers_global=ctx.world * M)
    cfg.update(over)
    spec = config.parse(cfg)
    eng = Engine(spec.to_struct())
    x0, extra = initial_state(name, spec, M, ctx.rank * M)
    eng.set_state(x0)
    if extra:
        eng.set_walker_bias(extra["k"], extra["r0"])
    return eng, spec, cfg


def bench_md(ctx, name, M, steps, warmup, precision="f32", sampler=None):
    """Device-resident timing of c2 / c3 / c4 / dump: K steps, CUDA events on the engine stream, L2 flushed
    between steps; max over ranks."""
    torch = ctx.torch
    eng, spec, cfg = make_engine(ctx, name, M, precision)
    md_steps = cfg["steps_per_launch"]
    eng.begin()

    dumping = bool(cfg.get("dump"))

    def drop_frames():
        if dumping:  # frames stay on the device; the ring is only rewound (no D2H in the timed region)
            eng.lib.nds_drain_frames(eng.h, None, 0, None)

    for _ in range(warmup):
        eng.run(md_steps)
        drop_frames()
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ctx.barrier()
    launches0 = eng.launch_count
    dev_ms = 0.0
    t_wall = time.perf_counter()
    for _ in range(steps):
        ctx.flush.zero_()
        torch.cuda.synchronize()
        eng.run(md_steps, sync=True)
        dev_ms += eng.last_run_ms()
        drop_frames()
    ctx.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    launches = eng.launch_count - launches0
    dev_ms, wall_ms = ctx.max_over_ranks(dev_ms, wall_ms)
    value = ctx.world * M * md_steps * steps / (dev_ms * 1e-3)
    rec = {"value": value, "unit": "walker-steps/s", "ms_per_step": dev_ms / steps, "steps": steps,
           "walkers_per_gpu": M, "md_steps_per_step": md_steps, "kernel": eng.kernel_name,
           "gpu_launches": launches, "wall_ms_total": wall_ms, "workload": WORKLOAD_TEXT[name]}
    if name == "c3":
        # count the steps that were actually integrated in the timed blocks (stopped walkers do nothing)
        basin, exit_step = eng.get_commit()
        lo = warmup * md_steps
        done = np.clip(exit_step, lo, None) - lo
        tot = ctx.sum_over_ranks(float(done.sum()), float((basin >= 0).sum()), float((basin == 1).sum()))
        rec["value"] = tot[0] / (dev_ms * 1e-3)
        rec["committor"] = {"shots": ctx.world * M, "committed": int(tot[1]), "to_basin_1": int(tot[2]),
                            "p_basin1_given_committed": tot[2] / max(tot[1], 1),
                            "integrated_fraction_of_nominal_steps": tot[0] / (ctx.world * M * md_steps * steps)}
    per_gpu = rec["value"] / ctx.world
    peaks = ctx.peaks
    if dumping:
        width = eng.frame_width
        frames_per_step = md_steps // cfg["dump_freq"]
        fbytes = 4 if eng.frames_f32 else 8
        nbytes = frames_per_step * width * min(cfg["dump_walkers"], M) * fbytes
        rec["dump_gb_per_s"] = nbytes / (dev_ms / steps * 1e-3) / 1e9
        rec["dump"] = {"walkers": min(cfg["dump_walkers"], M), "freq": cfg["dump_freq"], "fields": width,
                       "bytes_per_field": fbytes, "bytes_per_launch": nbytes,
                       "histogram": "48 x 40 bins, all walkers, every dump step" if cfg.get("hist_n") else None,
                       "note": "frames are recorded in a device ring and rewound between steps (no D2H in the timed region)"}
    if name in ("dump", "dump1"):
        width = eng.frame_width
        frames_per_step = md_steps // cfg["dump_freq"]
        fbytes = 4 if eng.frames_f32 else 8
        nbytes = frames_per_step * width * M * fbytes
        gbs = nbytes / (dev_ms / steps * 1e-3) / 1e9
        peak = peaks.get("hbm_gbs", 6650.0)
        rec["roofline"] = {
            "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
            "note": f"algorithmic dump bytes: {frames_per_step} frames x {width} {'fp32' if fbytes == 4 else 'fp64'} "
                    f"fields x {M} walkers per launch of {md_steps} steps; peak = MEASURED_PEAKS.json hbm_gbs "
                    f"({'measured' if 'hbm_gbs' in peaks else 'fallback'}); the launch also integrates the {md_steps} "
                    "steps (see compute_floor)"}
        if name in NCU_TRAFFIC and M == (1 << 20):
            rec["roofline"]["traffic"], rec["roofline"]["traffic_source"] = NCU_TRAFFIC[name]
    else:
        ach = per_gpu * FLOPS_PER_WALKER_STEP[name] / 1e12
        pk = ctx.fp32_peak_tflops()
        rec["roofline"] = {
            "bound": "fp32_fma", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk, "traffic": None,
            "note": f"algorithmic {FLOPS_PER_WALKER_STEP[name]:.0f} flop/walker-step (SURVEY 8d) x walker-steps/s per "
                    f"GPU; peak = {ctx.sm_count} SMs x 128 lanes x 2 x {ctx.sm_max_mhz:.0f} MHz (clocks.max.sm; "
                    "MEASURED_PEAKS.json holds no FP32 figure). HBM traffic is negligible (state is "
                    "register-resident for the whole launch)."}
        if name in NCU_TRAFFIC and M == (1 << 20):
            rec["roofline"]["traffic"], rec["roofline"]["traffic_source"] = NCU_TRAFFIC[name]
        if name in ISSUE_MODEL:
            model, src = ISSUE_MODEL[name]
            measured = 4 * ctx.sm_count * ctx.sm_max_mhz * 1e6 * 32.0 / per_gpu  # cycles per 32 walker-steps per SMSP
            rec["roofline"]["issue_slots"] = {"model_cycles_per_32_walker_steps": model, "measured_cycles": measured,
                                              "frac": model / measured, "source": src}
    return rec, eng, spec, cfg


def attach_exchange(ctx, eng, mode):
    if ctx.world == 1:
        return "none"
    from ndsimulator_b200 import distributed as D
    if mode == "p2p":
        D.attach_peer_hill_exchange(eng)
    elif mode == "nccl":
        D.attach_native_hill_exchange(eng)
    else:
        D.attach_hill_exchange(eng)  # torch.distributed callback (host-synchronised)
    return mode


def bench_c5(ctx, grid, exchange, strides, warmup, M=8192, precision="f32"):
    """C5: `strides` deposition strides of 100 MD steps.  Exact mode: the table grows by world * M hills per stride,
    the unit that matters is the walker-hill pair (one MUFU.EX2).  Grid mode: cost per step independent of the hills."""
    torch = ctx.torch
    name = "c5_grid" if grid else "c5"
    md_steps = 100
    total_strides = warmup + strides + (strides if grid else 0) + 2
    eng, spec, cfg = make_engine(ctx, name, M, precision, steps=md_steps * total_strides)
    mode = attach_exchange(ctx, eng, exchange)
    eng.begin()
    for _ in range(warmup):
        eng.run(md_steps)
    ctx.barrier()
    x0 = eng.exchange_info()
    launches0 = eng.launch_count
    dev_ms, pair_evals = 0.0, 0.0
    for _ in range(strides):
        ctx.flush.zero_()
        torch.cuda.synchronize()
        if not grid:
            pair_evals += M * md_steps * (eng.n_hills() + ctx.world * M)  # table after this stride's deposition
        eng.run(md_steps, sync=True)
        dev_ms += eng.last_run_ms()
    ctx.barrier()
    launches = eng.launch_count - launches0
    x1 = eng.exchange_info()
    xms = x1["exchange_ms"] - x0["exchange_ms"]
    dev_ms, xms = ctx.max_over_ranks(dev_ms, xms)
    value = ctx.world * M * md_steps * strides / (dev_ms * 1e-3)
    rec = {"value": value, "unit": "walker-steps/s", "ms_per_stride": dev_ms / strides, "strides": strides,
           "walkers_per_gpu": M, "md_steps_per_stride": md_steps, "kernel": eng.kernel_name,
           "hill_exchange": mode, "exchange_ms_per_stride": xms / strides, "exchange_share": xms / dev_ms,
           "gpu_launches": launches, "l2": "flushed between strides", "workload": WORKLOAD_TEXT[name]}
    if grid:
        # the same strides again, enqueued back to back in ONE nds_run (no host synchronisation between strides)
        ctx.barrier()
        eng.run(md_steps * strides, sync=True)
        (pipe_ms,) = ctx.max_over_ranks(eng.last_run_ms())
        rec["back_to_back"] = {"value": ctx.world * M * md_steps * strides / (pipe_ms * 1e-3),
                               "ms_per_stride": pipe_ms / strides,
                               "note": "all strides in one nds_run call, L2 not flushed (state + grid are < 2 MB)"}
        ach = value / ctx.world * GRID_FLOPS_PER_WALKER_STEP / 1e12
        rec["roofline"] = {"bound": "latency (8192 walkers = 16384 lanes of md_pair_grid = 512 warps on 592 "
                                    "schedulers: one warp each, bound by the length of a step's dependent "
                                    "instruction stream)", "achieved": ach,
                           "peak": ctx.fp32_peak_tflops(), "unit": "TFLOP/s", "frac": ach / ctx.fp32_peak_tflops(),
                           "traffic": None,
                           "note": f"{GRID_FLOPS_PER_WALKER_STEP:.0f} flop/walker-step (212 of SURVEY 8d + 97 of the "
                                   "bicubic interpolation); at 8192 walkers per GPU the launch is latency-bound, "
                                   "the figure of merit is walker-steps/s independent of the hill count"}
        rec["grid_points"] = [int(spec.mtd_grid_n[0]), int(spec.mtd_grid_n[1])]
        rec["evaluations_outside_grid"] = int(ctx.sum_over_ranks(float(eng.mtd_grid_outside()))[0])
    else:
        he = ctx.world * pair_evals / (dev_ms * 1e-3)
        mufu_peak = ctx.sm_count * 16 * ctx.sm_max_mhz * 1e6 * ctx.world
        rec["hill_evals_per_s"] = he
        rec["hills_final"] = eng.n_hills()
        rec["roofline"] = {"bound": "mufu_ex2", "achieved": he / 1e12, "peak": mufu_peak / 1e12,
                           "unit": "T hill-evals/s", "frac": he / mufu_peak, "traffic": None,
                           "note": "one MUFU.EX2 per walker-hill pair; peak = GPUs x SMs x 16 lanes x clocks.max.sm"}
    rec["exchange_check"] = exchange_check(ctx, eng, grid, M, md_steps)
    return rec


def exchange_check(ctx, eng, grid, M, md_steps):
    """In-run verification of the collective (N > 1): every rank must hold the same hills / the same grid, and the
    segment of one more deposition must be exactly the colvars all ranks held when it was deposited."""
    if ctx.world == 1:
        return "n/a (one rank)"
    eng.sync()
    if grid:
        g = eng.get_mtd_grid()
        digests = ctx.gather_objects(hashlib.sha256(np.ascontiguousarray(g).tobytes()).hexdigest())
        if len(set(digests)) != 1:
            return f"FAILED: {len(set(digests))} different grids over {ctx.world} ranks"
        if not (np.isfinite(g).all() and g[..., 0].max() > 0):
            return "FAILED: empty or non-finite grid"
        return "ok"
    n0 = eng.n_hills()
    sums = ctx.gather_objects(eng.table_checksum(0, n0))
    if len(set(sums)) != 1:
        return f"FAILED: table checksums differ over the ranks: {sums}"
    colv = ctx.gather_objects(eng.get_state(fields=("colv",))["colv"])  # pre-step colvars = the next hill centres
    eng.run(md_steps, sync=True)
    c, h = eng.get_hills()
    seg = c[n0:n0 + ctx.world * M]
    want = np.concatenate([cv.T for cv in colv])
    if seg.shape != want.shape or not np.array_equal(seg, want):
        return "FAILED: the new table segment is not the all-gathered colvars of the deposition step"
    sums = ctx.gather_objects(eng.table_checksum(0, eng.n_hills()))
    if len(set(sums)) != 1:
        return "FAILED: table checksums differ after the extra stride"
    return "ok"


def bench_e2e(ctx, eng, spec, cfg, name, M, md_steps, precision, steps):
    """End to end through the C ABI with pinned HOST buffers: one block per direction, every pass copies (x, v) H2D
    and (x, v, thermo) D2H (nds_submit / nds_wait).  value = ONE ensemble, its passes strictly in sequence; inside a
    pass the walkers are processed in chunks so that the copies of one chunk overlap the kernel of another."""
    torch = ctx.torch
    from ndsimulator_b200 import config
    from ndsimulator_b200.engine import Engine
    nd = spec.ndim
    f32 = precision == "f32"
    tdt, esz = (torch.float32, 4) if f32 else (torch.float64, 8)

    class Batch:
        def __init__(self, engine):
            self.eng, self.flip = engine, 0
            # ping-pong: the block read back from pass n is (its first 2 ndim rows) the input of pass n + 1
            self.blocks = [torch.empty((2 * nd + 5, M), dtype=tdt).pin_memory() for _ in range(2)]
            st = engine.get_state(fields=("x", "v"))
            self.blocks[0][:nd].copy_(torch.from_numpy(st["x"]).to(tdt))
            self.blocks[0][nd:2 * nd].copy_(torch.from_numpy(st["v"]).to(tdt))

        def submit(self, chunks):
            src, dst = self.blocks[self.flip], self.blocks[1 - self.flip]
            self.eng.submit(src.data_ptr(), md_steps, dst.data_ptr(), chunks)
            self.flip = 1 - self.flip

        def wait(self):
            self.eng.wait()
            if cfg.get("dump"):  # the frames of the pass stay on the device; rewind the ring
                self.eng.lib.nds_drain_frames(self.eng.h, None, 0, None)

    def timed(fn, n):
        ctx.barrier()
        t0 = time.perf_counter()
        fn(n)
        ctx.barrier()
        return ctx.max_over_ranks(time.perf_counter() - t0)[0]

    b0 = Batch(eng)

    def serial(chunks):
        def run(n):
            for _ in range(n):
                b0.submit(chunks)
                b0.wait()
        return run

    n = max(4, min(steps, 20))
    out = {"unit": "walker-steps/s", "h2d_bytes_per_step": 2 * nd * M * esz,
           "d2h_bytes_per_step": (2 * nd + 5) * M * esz, "steps": n}
    work = ctx.world * M * md_steps
    serial(1)(2)
    t1 = timed(serial(1), n)
    out["unchunked_value"] = work * n / t1
    chunks = int(os.environ.get("NDS_BENCH_CHUNKS", "8"))
    serial(chunks)(2)
    tc = timed(serial(chunks), n)
    out["value"] = work * n / tc
    out["mode"] = (f"one ensemble, passes in sequence; each pass = ONE pinned block H2D -> {md_steps} steps -> ONE pinned "
                   f"block D2H through nds_submit / nds_wait, pipelined over {chunks} walker chunks on separate copy / "
                   "compute streams; unchunked_value = the same with one copy, one launch, one copy in series")

    # ceiling of the host link: the same bytes, copies only (both directions at once), all ranks together
    def copies(nrep):
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        dev = torch.empty((2 * nd + 5, M), dtype=tdt, device="cuda")
        for _ in range(nrep):
            with torch.cuda.stream(s1):
                dev[:2 * nd].copy_(b0.blocks[0][:2 * nd], non_blocking=True)
            with torch.cuda.stream(s2):
                b0.blocks[1].copy_(dev, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
    copies(2)
    tcopy = timed(copies, n)
    out["copy_only_ms_per_pass"] = 1e3 * tcopy / n
    out["host_link_ceiling"] = work * n / max(tcopy, 1e-9)
    out["host_link_note"] = ("copy_only = the pass's H2D and D2H bytes moved concurrently with no kernel, all ranks at "
                             "once: walker-steps/s the host link alone would allow")
    if name in ("c2", "c2_nodump", "c4"):
        # second, independent ensemble (its own walker ids -> its own Philox streams): two passes in flight
        cfg2 = dict(cfg, walker_offset=(ctx.world + ctx.rank) * M, n_walkers_global=2 * ctx.world * M)
        eng2 = Engine(config.parse(cfg2).to_struct())
        x02, extra2 = initial_state(name, spec, M, (ctx.world + ctx.rank) * M)
        eng2.set_state(x02)
        if extra2:
            eng2.set_walker_bias(extra2["k"], extra2["r0"])
        eng2.begin()
        eng2.run(md_steps)
        if cfg.get("dump"):
            eng2.lib.nds_drain_frames(eng2.h, None, 0, None)
        b1 = Batch(eng2)
        bs = [b0, b1]

        def pipelined(n):
            bs[0].submit(1)
            for i in range(n):
                if i + 1 < n:
                    bs[(i + 1) % 2].submit(1)
                bs[i % 2].wait()
        pipelined(4)
        tp = timed(pipelined, 2 * n)
        out["two_ensembles_value"] = work * 2 * n / tp
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=["c2", "c2_nodump", "c3", "c4", "c5", "c5_grid", "dump", "dump1"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: per workload)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl", "callback"],
                    help="C5 exchange: peer-memory stores + device barrier, native NCCL, or the torch callback")
    ap.add_argument("--strides", type=int, default=30, help="C5: deposition strides in the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="headline workload only")
    ap.add_argument("--port-reference", action="store_true", help="--impl reference: time the C port, not Python")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    if args.impl == "reference":
        reference_arm(args)
        return

    ctx = Ctx(args)
    name = args.workload
    world, rank = ctx.world, ctx.rank
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    default_M = {"c2": 1 << 20, "c2_nodump": 1 << 20, "c3": 1 << 20, "c4": (1 << 20) // world, "c5": 8192,
                 "c5_grid": 8192, "dump": 1 << 20, "dump1": 1 << 20}
    M = args.walkers or default_M[name]

    if name in ("c5", "c5_grid"):
        if sampler:
            sampler.start()
        rec = bench_c5(ctx, name == "c5_grid", args.exchange, args.strides, args.warmup, M, args.precision)
        head = {"value": rec["value"], "ms_per_step": rec["ms_per_stride"], "steps": rec["strides"],
                "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"],
                "config": {k: v for k, v in rec.items() if k not in ("value", "roofline", "gpu_launches")}}
        e2e = None
    else:
        if name == "c3":
            args.steps = 10 - args.warmup  # the whole job is 10 blocks of 1000 steps (steps: 10000)
        rec, eng, spec, cfg = bench_md(ctx, name, M, args.steps, args.warmup, args.precision, sampler)
        head = {"value": rec["value"], "ms_per_step": rec["ms_per_step"], "steps": args.steps,
                "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"],
                "config": {"workload": rec["workload"], "walkers_per_gpu": M,
                           "md_steps_per_step": rec["md_steps_per_step"], "kernel": rec["kernel"],
                           "l2": "flushed between steps (256 MiB memset)",
                           "timing": "CUDA events on the engine stream per step, summed; max over ranks",
                           "wall_ms_total": rec["wall_ms_total"]}}
        if "committor" in rec:
            head["committor"] = rec["committor"]
        e2e = None
        if "dump" in rec:
            head["config"]["dump"] = rec["dump"]
            head["config"]["dump_gb_per_s"] = rec["dump_gb_per_s"]
        if not args.no_e2e and name in ("c2", "c2_nodump", "c4"):
            e2e = bench_e2e(ctx, eng, spec, cfg, name, M, rec["md_steps_per_step"], args.precision, args.steps)
        del eng
    clocks = sampler.stop() if sampler else None

    workloads = None
    if name == "c2" and not args.no_workloads:
        workloads = {}
        w3, *_ = bench_md(ctx, "c3", 1 << 20, 7, 3)
        workloads["c3"] = w3
        w4, *_ = bench_md(ctx, "c4", (1 << 20) // world, 5, 3)
        w4["scaling"] = "strong: the 256 windows x 4096 walkers are sharded over the ranks"
        workloads["c4"] = w4
        wn, *_ = bench_md(ctx, "c2_nodump", 1 << 20, 10, 3)
        workloads["c2_nodump"] = wn
        for dn in ("dump", "dump1"):
            wd, *_ = bench_md(ctx, dn, 1 << 20, 10, 3)
            # what the same launch costs without frames: the part of the time the dump bytes cannot be blamed for
            wd["compute_floor_ms"] = wn["ms_per_step"] * wd["md_steps_per_step"] / wn["md_steps_per_step"]
            workloads[dn] = wd
        workloads["c5"] = bench_c5(ctx, False, "p2p", args.strides, 3)
        if world > 1:
            workloads["c5_nccl"] = bench_c5(ctx, False, "nccl", args.strides, 3)
        workloads["c5_grid"] = bench_c5(ctx, True, "p2p", 100, 3)
        if world > 1:
            workloads["c5_grid_nccl"] = bench_c5(ctx, True, "nccl", 100, 3)

    if rank == 0:
        out = {
            "metric": "walker-steps/s", "value": head["value"], "unit": "walker-steps/s", "n_gpus": world,
            "steps": head["steps"], "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic", "config": head["config"],
            "gpu_launches": head["gpu_launches"], "clocks": clocks, "roofline": head["roofline"],
        }
        if "committor" in head:
            out["committor"] = head["committor"]
        if e2e:
            out["e2e"] = e2e
        if workloads:
            out["workloads"] = workloads
            out["gpu_launches"] += sum(int(w.get("gpu_launches", 0)) for w in workloads.values())
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(name if name in ("c3", "c4", "c5") else "c2_nodump")
        print(json.dumps(out))
    if world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
