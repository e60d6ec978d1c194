"""``NDRun``: the reference's orchestrator interface (``ndsimulator/ndrun.py:51-268``) in front of the
B200 multi-walker engine.

Same constructor vocabulary (every YAML key of the reference that reaches the MD / committor path,
SURVEY.md section 8b), same ``begin() / run() / end()`` protocol, same log / dump files -- but the loop
body runs as fused CUDA launches over ``n_walkers`` independent walkers instead of one Python iteration
per step.  New keys: ``n_walkers``, ``precision`` (f32|f64), ``steps_per_launch``, ``device``,
``dump_walkers``, ``mtd_shared``, ``umb_per_walker``.

    from ndsimulator_b200 import NDRun
    sim = NDRun(ndim=2, potential="Mueller2d", x0=[22.0, 24.0], mass=1.0, gamma=0.1, dt=1.0,
                steps=100000, n_walkers=1 << 20, precision="f32", dump=False)
    sim.begin(); sim.run(); sim.end()
    sim.positions        # [ndim][n_walkers]
"""
import logging
import os

import numpy as np

from . import config
from .constant import kB
from .dump import DumpWriter
from .engine import Engine


class _Atoms:
    """``run.atoms`` of the reference (data.py:52-81) as a view on one walker of the engine state."""

    def __init__(self, run, walker=0):
        self._run = run
        self._w = walker

    def _get(self, key):
        st = self._run._state()
        v = st[key]
        return v[..., self._w].copy() if v.ndim > 1 else v[self._w]

    positions = property(lambda self: self._get("x"))
    velocities = property(lambda self: self._get("v"))
    forces = property(lambda self: self._get("f"))
    fixf = property(lambda self: self._get("fixf"))
    colv = property(lambda self: self._get("colv"))
    pe = property(lambda self: float(self._get("pe")))
    ke = property(lambda self: float(self._get("ke")))
    biase = property(lambda self: float(self._get("biase")))
    totale = property(lambda self: float(self._get("totale")))
    T = property(lambda self: float(self._get("T")))


class _Modify:
    """``run.modify`` (engines/modify.py:38-63): state setters applied to every walker."""

    def __init__(self, run):
        self._run = run

    def set_positions(self, x):
        run = self._run
        x = np.asarray(x, dtype=float)
        if x.ndim == 1:
            x = np.repeat(x.reshape(-1, 1), run.n_walkers, axis=1)
        run._x0 = x
        run._push_state()

    def set_velocity(self, T=None, v=None):
        run = self._run
        if T is not None:
            run.spec.initial_temperature = float(T)
            run._v0 = None
            run._rebuild()
        elif v is not None:
            v = np.asarray(v, dtype=float)
            if v.ndim == 1:
                v = np.repeat(v.reshape(-1, 1), run.n_walkers, axis=1)
            run._v0 = v
            run._push_state()
        else:
            raise RuntimeError("velocity need to be assigned")


class NDRun:
    def __init__(self, ndim, **kwargs):
        all_kwargs = dict(kwargs, ndim=ndim)
        self.kwargs = all_kwargs
        self.spec = config.parse(all_kwargs)
        spec = self.spec
        for key in ("ndim", "temperature", "mass", "steps", "method", "seed", "screen", "dump"):
            setattr(self, key, getattr(spec, key))
        self.biases = list(spec.biases)
        self.initial_temperature = spec.initial_temperature
        self.n_walkers = spec.n_walkers
        self.x0 = spec.x0
        self._x0 = np.repeat(spec.x0.reshape(-1, 1), spec.n_walkers, axis=1)
        spread = float(all_kwargs.get("x0_spread", 0.0) or 0.0)
        if spread > 0:  # new key: walkers start from a Gaussian cloud around x0 (seeded, rank independent)
            gid = spec.walker_offset + np.arange(spec.n_walkers)
            rs = np.random.RandomState(spec.seed & 0x7FFFFFFF)
            cloud = rs.normal(scale=spread, size=(spec.ndim, spec.n_walkers_global or spec.n_walkers))
            self._x0 = self._x0 + cloud[:, gid]
        self._v0 = None
        self._window = None
        self.engine = Engine(spec.to_struct())
        if spec.path_ref is not None:
            self.engine.set_path(spec.path_ref, spec.path_ind, spec.path_sigma)
        grid = all_kwargs.get("umb_window_grid")
        if grid is not None:  # new key: one umbrella window per walker (walker w -> window w mod n_windows)
            axes = [np.linspace(lo, hi, int(n)) for lo, hi, n in grid]
            mesh = np.stack([m.reshape(-1) for m in np.meshgrid(*axes, indexing="xy")])
            gid = spec.walker_offset + np.arange(spec.n_walkers)
            r0 = mesh[:, gid % mesh.shape[1]]
            k = np.repeat(np.asarray(spec.umb_k[0], dtype=float).reshape(-1, 1), spec.n_walkers, axis=1)
            self._window = (k, r0)
            self.engine.set_walker_bias(k, r0)
            if spec.colvar == "Proj5dto2d":  # start on the window centre: xi1 = |(x0,x1)|, xi2 = |(x2,x3)|
                self._x0 = np.stack([r0[0], np.zeros_like(r0[0]), r0[1], np.zeros_like(r0[0]), self._x0[4]])
            elif spec.colvar == "Original":
                self._x0 = r0.copy()
        self.engine.set_state(self._x0)
        self.atoms = _Atoms(self)
        self.modify = _Modify(self)
        self.step = 0
        self.time = 0.0
        self.commit_basin = -1
        self._cache = None
        self._dump = None
        workdir = os.path.join(spec.root, spec.run_name)
        os.makedirs(workdir, exist_ok=True)
        self.logfile = os.path.join(workdir, "log")
        self._logger = logging.getLogger(self.logfile)  # ndrun.py:105-107
        self._logger.propagate = False
        self._logger.setLevel(logging.INFO)
        for h in list(self._logger.handlers):
            self._logger.removeHandler(h)
        fh = logging.FileHandler(self.logfile, mode="w")
        fh.setFormatter(logging.Formatter("%(message)s"))
        self._logger.addHandler(fh)
        if spec.screen:
            sh = logging.StreamHandler()
            sh.setFormatter(logging.Formatter("%(message)s"))
            self._logger.addHandler(sh)

    # ------------------------------------------------------------------ helpers
    @property
    def logger(self):
        return self._logger

    def _state(self):
        if self._cache is None:
            self._cache = self.engine.get_state()
        return self._cache

    def _push_state(self):
        self.engine.set_state(self._x0, self._v0)
        self._cache = None

    def _rebuild(self):
        self.engine.close()
        self.engine = Engine(self.spec.to_struct())
        if self.spec.path_ref is not None:
            self.engine.set_path(self.spec.path_ref, self.spec.path_ind, self.spec.path_sigma)
        self.engine.set_state(self._x0, self._v0)
        if self._window is not None:
            self.engine.set_walker_bias(*self._window)
        self._cache = None

    def set_seed(self, seed):  # ndrun.py:173-176
        self.spec.seed = int(seed)
        self.seed = int(seed)
        self._rebuild()

    def set_umbrella_windows(self, k, r0):
        """Per-walker umbrella windows (k, r0 of shape [colvardim][n_walkers]); needs umb_per_walker."""
        self._window = (np.asarray(k, dtype=float), np.asarray(r0, dtype=float))
        self.engine.set_walker_bias(*self._window)

    # state of all walkers
    positions = property(lambda self: self._state()["x"])
    velocities = property(lambda self: self._state()["v"])
    colv = property(lambda self: self._state()["colv"])
    temperatures = property(lambda self: self._state()["T"])

    # ------------------------------------------------------------------ protocol
    def begin(self):  # ndrun.py:182-206
        spec = self.spec
        self.engine.begin()
        self.step, self.time = 0, 0.0
        self._cache = None
        self._dep_steps = []
        if spec.dump:
            self._dump = DumpWriter(spec, int(min(spec.dump_walkers, spec.n_walkers)))
            self._dump.open_files()
            if "umb" in spec.biases:  # Harmonic.VR after MD.begin's no-arg compute (umb.py:32-33)
                c = self._state()["colv"][:, :self._dump.n]
                vr = np.zeros(self._dump.n)
                if self._window is not None:
                    k, r0 = self._window[0][:, :self._dump.n], self._window[1][:, :self._dump.n]
                else:
                    k, r0 = spec.umb_k[0].reshape(-1, 1), spec.umb_r0[0].reshape(-1, 1)
                vr = 0.5 * np.sum(k * (c - r0) ** 2, axis=0)
                self._dump.umb_vr = vr
            self._flush_frames()

    def _hills_at(self, step):
        per = self.spec.n_walkers_global if self.spec.mtd_shared else 1
        return per * sum(1 for s in self._dep_steps if s < step)

    def _replay_deposits(self, n):
        """Steps at which Gaussian.update deposits (gaus.py:134-136), replayed with the same fp64 clock."""
        spec = self.spec
        if "mtd" not in spec.biases:
            return
        t, s = self.time, self.step
        count = len(self._dep_steps)
        for _ in range(n):
            if int(np.floor(t / spec.mtd_dep_freq)) + 1 > count:
                self._dep_steps.append(s)
                count += 1
            t += spec.dt
            s += 1

    def _flush_frames(self):
        if not self._dump:
            return
        frames = self.engine.drain_frames()
        if len(frames) == 0:
            return
        hills = self.engine.get_hills(0) if "mtd" in self.spec.biases else None
        if hills is not None and (self.spec.mtd_adapsig_geo or self.spec.mtd_adapsig_hist):
            hills = hills + (self.engine.get_hill_sigma2(0),)
        self._dump.write_frames(frames, self.engine.frame_fields(), hills=hills, hills_at=self._hills_at)
        if self.spec.screen:  # ndrun.py:222-224: "<step> temperature: <T>" every dump.freq steps
            names = self.engine.frame_fields()
            for fr in frames:
                s = int(fr[names.index("step"), 0])
                if s % self.spec.dump_freq == 0 and s < self.spec.steps:
                    self.logger.info(f"{s} temperature: {fr[names.index('T'), 0]}")

    def run(self, steps=None):  # ndrun.py:214-268
        spec = self.spec
        total = spec.steps if steps is None else steps
        if spec.method == "minimize":
            # Minimize.update returns None, so the reference's loop `while step < steps and nostop`
            # (ndrun.py:220,230) ends after ONE update; kept as is
            total = min(total, self.step + 1)
        nostop = True
        chunk = max(spec.steps_per_launch, 1)
        if spec.dump:
            per = max(spec.dump_freq, 1)
            cap = max((spec.dump_capacity - 2) // 2, 1) * per
            chunk = max(min(chunk * 8, cap), 1)
        else:
            chunk = chunk * 64
        while self.step < total and nostop:
            n = min(chunk, total - self.step)
            self._replay_deposits(n)
            self.engine.run(n)
            self.step = self.engine.step
            self.time = self.engine.time
            self._cache = None
            self._flush_frames()
            if spec.method == "committor" and self.engine.count_alive() == 0:
                nostop = False
        if spec.method == "committor":
            basin, exit_step = self.engine.get_commit()
            self.commit_basins, self.exit_steps = basin, exit_step
            self.commit_basin = int(basin[0])  # ndrun.py:264
            nostop = bool(np.any(basin < 0))
            if self.n_walkers == 1:
                self.step = int(exit_step[0])
        if spec.method in ("mc", "wl-mc"):  # ndrun.py:235-238,265-268: time advances on accepted moves only
            self.accepted = self.engine.get_mc()
            self.accept_rate = float(self.accepted[0])
            self.time = float(self.accepted[0]) * spec.dt
        self.logger.info(f"end condition {self.step} {total} {nostop}")
        if spec.method in ("mc", "wl-mc") and spec.screen:
            self.logger.info(f"!! MC acceptance ratio: {self.accept_rate / max(total, 1)}")
        if spec.method == "committor" and spec.screen:
            self.logger.info(f"!! Commit to basin {self.commit_basin}")
            if self.n_walkers > 1:
                counts = {int(b): int((basin == b).sum()) for b in np.unique(basin)}
                self.logger.info(f"!! commit counts over {self.n_walkers} walkers: {counts}")

    def end(self):  # ndrun.py:208-212
        if self._dump:
            self._dump.close_files()
            self._dump = None
        for h in list(self._logger.handlers):
            h.flush()

    # ------------------------------------------------------------------ observables
    def free_energy_grid(self, boundary=((0.0, 48.0), (0.0, 40.0)), increment=(1.0, 1.0), walker=0):
        """Sum of hills on ``arange(b0, b1, inc)`` per CV: the reference's only free-energy estimate
        (bias/gaus.py:365-393 on the grid of io/plot.py:517-519)."""
        nx = int(np.ceil((boundary[0][1] - boundary[0][0]) / increment[0]))
        if self.spec.colvardim == 1:
            return self.engine.fes_grid(boundary[0][0], increment[0], nx, walker=walker)[0]
        ny = int(np.ceil((boundary[1][1] - boundary[1][0]) / increment[1]))
        return self.engine.fes_grid(boundary[0][0], increment[0], nx, boundary[1][0], increment[1], ny,
                                    walker=walker)

    def colvar_histogram(self, boundary=((0.0, 48.0), (0.0, 40.0)), increment=(1.0, 1.0)):
        nx = int(np.ceil((boundary[0][1] - boundary[0][0]) / increment[0]))
        if self.spec.colvardim == 1:
            return self.engine.colvar_histogram(boundary[0][0], increment[0], nx)[0]
        ny = int(np.ceil((boundary[1][1] - boundary[1][0]) / increment[1]))
        return self.engine.colvar_histogram(boundary[0][0], increment[0], nx, boundary[1][0], increment[1], ny)

    def save_checkpoint(self, path):
        """(x, v, clock, hill table, VR) and the host-side dump bookkeeping -> .npz; with the same config + seed the
        run -- and its dump files -- continue exactly where they stopped."""
        ck = self.engine.checkpoint()
        ck["host_dep_steps"] = np.asarray(getattr(self, "_dep_steps", []), dtype=np.int64)
        if self._dump is not None:
            ck["host_hills_dumped"] = self._dump.hills_dumped
            if self._dump.lag is not None:
                ck["host_lag"] = np.asarray(self._dump.lag)
            if self._dump.umb_vr is not None:
                ck["host_umb_vr"] = np.asarray(self._dump.umb_vr)
        np.savez(path, **{k: np.asarray(v) for k, v in ck.items()})

    def load_checkpoint(self, path):
        z = np.load(path)
        ck = {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}
        self.engine.restore(ck)
        self.step, self.time = int(ck["step"]), float(ck["time"])
        # deposition steps so far: fix{i}.dat lists the hills that existed at each dump row (_hills_at)
        self._dep_steps = [int(s) for s in np.atleast_1d(ck.get("host_dep_steps", []))]
        self._cache = None
        if self.spec.dump:
            # Engine.restore() ran begin(), which recorded a frame labelled step 0: not part of a continued run
            self.engine.discard_frames()
            if self._dump is None:  # fresh process: continue the files of the interrupted run
                self._dump = DumpWriter(self.spec, int(min(self.spec.dump_walkers, self.spec.n_walkers)))
                self._dump.open_files(append=True)
            self._dump.hills_dumped = int(ck.get("host_hills_dumped", 0))
            self._dump.lag = ck.get("host_lag")
            self._dump.umb_vr = ck.get("host_umb_vr")

    @classmethod
    def from_dict(cls, dictionary):  # ndrun.py:270-282
        return cls(**dict(dictionary))

    def as_dict(self):
        return dict(self.kwargs)


def kinetic_temperature(v, m, ndim):
    """T = ke / ndim / kB * 2 (md.py:227-228) for host-side post-processing."""
    return np.sum(v * v, axis=0) * m / 2.0 / ndim / kB * 2.0
