"""Reference-side engine plugin: the B200 multi-walker engine behind the reference's OWN ``NDRun``.

The reference builds its engine in ``engine_from_config`` (``ndsimulator/engines/__init__.py:11-31``) and then talks
to it through four members only -- ``initialize(run)``, ``begin()``, ``update(step, time) -> nostop`` and
``current_dt`` (``ndrun.py:152,186,230,236-240``).  :class:`B200Engine` implements exactly that protocol on top of
``libndsb200.so``.  The three lines a maintainer adds to the factory::

    elif method == "b200":
        from ndsimulator_b200.plugin import B200Engine
        return B200Engine(config)

With ``method: b200`` the UNMODIFIED ``NDRun.run`` loop, ``Stat``, ``Dump`` and ``Modify`` keep working: they see
walker 0 through ``run.atoms`` exactly as with ``method: md``, while ``n_walkers`` walkers advance on the GPU.
Because the reference's loop calls ``update`` once per step, the plugin runs the fused kernel a block of
``steps_per_launch`` steps AHEAD, records one frame of walker 0 per step on the device, and hands the frames out one
call at a time.  Potential, colvar and biases are evaluated inside the kernel; the reference-side ``fix`` objects
are replaced by proxies that only serve ``Dump`` (``fix.VR``, ``fix.dump_data()``; ``io/dump.py:98-101``).

New config keys (all optional): ``n_walkers`` (1), ``precision`` ("f64"), ``steps_per_launch`` (256),
``b200_method`` ("md": the engine method the block runs).  Everything else is the reference's vocabulary.
"""
import numpy as np

from . import config as nds_config
from .engine import Engine


class _FixProxy:
    """Stands where the reference keeps a ``Harmonic`` / ``Gaussian``: ``NDRun.run`` calls ``update`` (a no-op here,
    the kernel deposits), ``Dump.write`` reads ``VR`` and ``dump_data()``."""

    def __init__(self, owner, kind):
        self.owner, self.kind = owner, kind
        self.VR = 0
        self._dumped = 0

    def initialize(self, run):
        pass

    def update(self, step, time):
        pass

    def dump_data(self):
        if self.kind != "mtd":
            return None  # umb.py:75-76 returns nothing: Dump prints "None"
        o = self.owner
        c, h = o.engine.get_hills(0)
        s2 = o.engine.get_hill_sigma2(0) if o.spec.mtd_adapsig_geo or o.spec.mtd_adapsig_hist else None
        upto = min(len(h), o.hills_at(o.run.step))
        line = ""
        for i in range(self._dumped, upto):  # gaus.py:440-461
            line += str(i) + " " + " ".join(map(str, c[i, :])) + " "
            if s2 is not None:
                line += str(s2[i]) + " "
            else:
                line += " ".join(map(str, np.asarray(o.spec.mtd_sigma, dtype=float).flat)) + " "
            line += str(h[i]) + "\n"
        self._dumped = upto
        return line


class B200Engine:
    def __init__(self, config):
        self.config = dict(config)
        self.current_dt = None
        self.commit_basin = -1
        self.accept = True

    # ---- AllData protocol (data.py:17-49): the members other components read from the engine
    def initialize(self, run):
        cfg = dict(self.config)
        self.run = run
        self.atoms = run.atoms
        self.block = int(cfg.pop("steps_per_launch", 256))
        cfg["method"] = cfg.pop("b200_method", "md")
        cfg.setdefault("precision", "f64")
        cfg.setdefault("n_walkers", 1)
        # one frame of walker 0 after EVERY step: what NDRun / Stat / Dump read between two update() calls
        cfg.update(dump=True, dump_freq=1, stat_freq=1, dump_walkers=1, dump_capacity=self.block + 4,
                   steps_per_launch=self.block, screen=False)
        self.spec = nds_config.parse(cfg)
        self.engine = Engine(self.spec.to_struct())
        if self.spec.path_ref is not None:
            self.engine.set_path(self.spec.path_ref, self.spec.path_ind, self.spec.path_sigma)
        self.current_dt = self.spec.dt
        self.names = self.engine.frame_fields()
        self.idx = {n: i for i, n in enumerate(self.names)}
        self.frames = {}
        self.started = False
        self.dep_steps = []
        # the kernel evaluates the biases: the reference-side fixes become dump-only proxies
        kinds = []
        for b in self.spec.biases:
            kinds += ["umb"] * max(len(self.spec.umb_k), 1) if b == "umb" else [b]
        run.fixes[:] = [_FixProxy(self, k) for k in kinds]
        self.fixes = run.fixes

    def begin(self):  # MD.begin, md.py:101-110: energies and forces at the initial point
        a = self.atoms
        ev = self.engine.eval_at(np.asarray(a.positions, dtype=float).reshape(-1, 1))
        a.pe = float(ev["V"][0])
        a.forces = ev["F"][:, 0].copy()
        a.biase = float(ev["Vb"][0]) if self.spec.biases else 0
        a.fixf = ev["Fb"][:, 0].copy()
        np.copyto(a.colv, np.reshape(ev["colv"][:, 0], np.shape(a.colv)))  # keeps the (1, 1) shape of two.py (A28)
        a.totale = a.ke + a.pe + a.biase
        umb = [fx for fx in self.fixes if fx.kind == "umb"]
        if len(umb) == 1:  # Harmonic.VR after the no-argument compute of MD.begin (umb.py:32-33); hills: none yet
            k = np.asarray(self.spec.umb_k[0], dtype=float).reshape(-1)
            r0 = np.asarray(self.spec.umb_r0[0], dtype=float).reshape(-1)
            umb[0].VR = float(0.5 * np.sum(k * (ev["colv"][:, 0] - r0) ** 2))
        self.started = False

    def hills_at(self, step):
        per = self.spec.n_walkers_global if self.spec.mtd_shared else 1
        return per * sum(1 for s in self.dep_steps if s < step)

    def _start(self):
        """First update(): NDRun.begin has set the velocities by now (ndrun.py:191-192)."""
        a, M = self.atoms, self.spec.n_walkers
        x = np.repeat(np.asarray(a.positions, dtype=float).reshape(-1, 1), M, axis=1)
        v = None
        if a.velocities is not None:
            scale = np.sqrt(8.6173303e-5 * self.run.initial_temperature / (104.23315626367025 * self.run.mass))
            v = np.random.RandomState(int(self.spec.seed) & 0x7FFFFFFF).normal(scale=scale, size=x.shape)
            v[:, 0] = a.velocities  # walker 0 is the reference's own trajectory
        self.engine.set_state(x, v)
        self.engine.begin()
        self.engine.discard_frames()
        self.started = True

    def _run_block(self, step, time):
        n = max(1, min(self.block, int(self.run.steps) - step))
        if "mtd" in self.spec.biases:  # Gaussian.update's schedule (gaus.py:134-136) on the same fp64 clock
            t, count = time, len(self.dep_steps)
            for k in range(n):
                if int(np.floor(t / self.spec.mtd_dep_freq)) + 1 > count:
                    self.dep_steps.append(step + k)
                    count += 1
                t += self.spec.dt
        self.engine.run(n)
        fr = self.engine.drain_frames()
        self.frames = {int(f[self.idx["step"], 0]): f[:, 0] for f in fr}

    def update(self, step, time):  # engine.update, ndrun.py:230: one step of walker 0 becomes visible
        if not self.started:
            self._start()
        if step + 1 not in self.frames:
            self._run_block(step, time)
        f, ix, a, nd, cd = self.frames.pop(step + 1), self.idx, self.atoms, self.spec.ndim, self.spec.colvardim
        np.copyto(a.prev_positions, a.positions)
        np.copyto(a.positions, [f[ix[f"x{d}"]] for d in range(nd)])
        np.copyto(a.colv, np.reshape([f[ix[f"c{k}"]] for k in range(cd)], np.shape(a.colv)))
        np.copyto(a.velocities, [f[ix[f"v{d}"]] for d in range(nd)])
        a.forces = np.array([f[ix[f"f{d}"]] for d in range(nd)])
        a.pe, a.ke, a.T = float(f[ix["pe"]]), float(f[ix["ke"]]), float(f[ix["T"]])
        a.biase = float(f[ix["biase"]]) if self.spec.biases else 0
        a.totale = float(f[ix["totale"]])
        n_umb = 0
        for fx in self.fixes:
            if fx.kind == "umb":  # refreshed by every deposition when a metadynamics bias is present (gaus.py:138-146)
                if f"VRu{n_umb}" in ix:
                    fx.VR = float(f[ix[f"VRu{n_umb}"]])
                n_umb += 1
            if fx.kind == "mtd":
                fx.VR = float(f[ix["VR"]])
        self.current_dt = self.spec.dt
        return True

    # ---- all walkers
    def get_state(self):
        """State of every walker at the END of the block the kernel has run (walker 0 may be ahead of run.atoms)."""
        return self.engine.get_state()


def register(engines_module=None):
    """Apply the three-line patch of INTEGRATION.md section 1 at run time: ``method: b200`` in the reference's
    ``engine_from_config``.  ``engines_module`` defaults to the importable ``ndsimulator.engines``; ``ndsimulator.ndrun``
    (which imported the factory by name) is patched too."""
    import importlib
    if engines_module is None:
        engines_module = importlib.import_module("ndsimulator.engines")
    original = engines_module.engine_from_config
    if getattr(original, "_b200", False):
        return original

    def engine_from_config(method, config):
        if method == "b200":
            return B200Engine(config)
        return original(method, config)

    engine_from_config._b200 = True
    engines_module.engine_from_config = engine_from_config
    try:
        ndrun = importlib.import_module("ndsimulator.ndrun")
        ndrun.engine_from_config = engine_from_config
    except ImportError:
        pass
    return engine_from_config
