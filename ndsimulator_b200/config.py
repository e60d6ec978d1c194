"""YAML / kwargs vocabulary of the reference -> ``nds_config``.

Mirrors how the reference builds its plugins: every component is constructed by
``pyfile_utils.instantiate(cls, prefix=..., optional_args=config)`` which fills a constructor
parameter ``p`` from the bare key ``p`` and then from ``<prefix>_p`` for each prefix in order, later
prefixes winning; unknown keys are silently ignored (reference ``ndrun.py:105-146``,
``engines/__init__.py:30``, ``bias/__init__.py:28-36``, ``potentials/_build.py:10-20``,
``colvars/_build.py:10-22``).  Class names resolve like
``potentials/_build.py:12-18``: a bare built-in name or ``ndsimulator.<pkg>.<name>``.

User-defined Python potentials / colvars (dotted paths or callables, which the reference would import
and call on the host every step) are REJECTED: this engine has no host fallback.
"""
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _abi
from ._abi import NdsConfig

_MISSING = object()


def pick(config: dict, name: str, prefixes, default=_MISSING):
    """Value a constructor parameter ``name`` receives under the reference's prefix rules."""
    if isinstance(prefixes, str):
        prefixes = [prefixes]
    val = config.get(name, _MISSING)
    for p in prefixes:
        key = f"{p}_{name}"
        if key in config:
            val = config[key]
    if val is _MISSING:
        return default
    return val


def _builtin_name(name, package: str, table: dict, what: str) -> str:
    if not isinstance(name, str):
        raise ValueError(
            f"user-defined Python {what} {name!r} cannot run on the B200 engine (no CPU fallback); "
            f"built-ins: {sorted(table)}")
    if what == "colvar" and name.rsplit(".", 1)[-1] == "Map50to2d":
        # colvars/misc.py:19-156: ndim = 50, and no potential of the reference accepts ndim = 50 (potential.py:12
        # with mueller.py:59 / misc.py:53), so no reference run can use it either
        raise ValueError("colvar Map50to2d needs ndim = 50, which no potential accepts (neither here nor in the "
                         "reference); its values and Jacobians are served by ndsimulator_b200.map50to2d(x) / "
                         "nds_map50to2d()")
    short = name
    for pre in (f"ndsimulator.{package}.", f"ndsimulator_b200.{package}."):
        if short.startswith(pre):
            short = short[len(pre):]
    if short.startswith("ndsimulator.") and short.rsplit(".", 1)[-1] in table:
        short = short.rsplit(".", 1)[-1]  # e.g. ndsimulator.potentials.mueller.Mueller2d
    if short not in table:
        if "." in name:
            raise ValueError(
                f"user-defined Python {what} '{name}' cannot run on the B200 engine: every step would "
                f"call back into host Python. Built-ins: {sorted(table)}")
        raise ValueError(f"cannot load the {what} {name}")  # potentials/_build.py:18
    return short


@dataclass
class RunSpec:
    """Parsed run description (host side)."""
    ndim: int
    temperature: float = 300.0
    initial_temperature: float = 300.0
    mass: float = 5.0
    steps: int = 100
    method: str = "md"
    seed: int = 0
    x0: np.ndarray = None
    potential: str = "Gaussian"
    pot_params: List[float] = field(default_factory=lambda: [0.0] * 16)
    colvar: str = "Original"
    true_colvar: str = "Original"
    colvardim: int = 2
    dt: float = 0.5
    gamma: float = 0.005
    integrate: str = "langevin"
    rescale_freq: Optional[float] = None
    biases: List[str] = field(default_factory=list)
    umb_k: List[np.ndarray] = field(default_factory=list)
    umb_r0: List[np.ndarray] = field(default_factory=list)
    umb_per_walker: bool = False
    path_ref: Optional[np.ndarray] = None     # Path / Path5dto2d colvar (colvars/path.py:32-71)
    path_ind: Optional[np.ndarray] = None
    path_sigma: float = 1.0
    mtd_w: float = 0.0
    mtd_sigma: Optional[np.ndarray] = None
    mtd_dep_freq: float = 10
    mtd_biasf: float = 1.0
    mtd_shared: bool = True
    mtd_capacity: int = 0
    mtd_adapsig_geo: bool = False
    mtd_adapsig_hist: bool = False
    mtd_grid: bool = False
    mtd_grid_lo: Optional[np.ndarray] = None
    mtd_grid_hi: Optional[np.ndarray] = None
    mtd_grid_n: Optional[np.ndarray] = None
    hist_n: Optional[np.ndarray] = None        # in-kernel colvar histogram (new keys hist_n / hist_lo / hist_hi)
    hist_lo: Optional[np.ndarray] = None
    hist_hi: Optional[np.ndarray] = None
    pe_thred: float = 0.0
    n_basins: int = 0
    criteria: Optional[np.ndarray] = None
    diffusion_time: int = 0
    mc_bounds: Optional[np.ndarray] = None
    mc_global_bounds: Optional[np.ndarray] = None
    mc_propose_dim: str = "single"
    mc_bound_cond: str = "periodic"
    emin: float = -1.2
    min_gtol: float = 1.0e-5
    min_maxiter: Optional[int] = None
    dump: bool = True
    dump_freq: int = 1
    stat_freq: int = 1
    dump_walkers: int = 1
    dump_capacity: int = 0
    screen: bool = True
    root: str = "./"
    run_name: str = "ndrun"
    # new keys
    n_walkers: int = 1
    walker_offset: int = 0
    n_walkers_global: int = 0
    precision: str = "f64"
    steps_per_launch: int = 1000
    device: int = 0
    noise_mode: int = _abi.NDS_NOISE_PHILOX

    def to_struct(self) -> NdsConfig:
        c = NdsConfig()
        c.abi_version = _abi.NDS_ABI_VERSION
        c.device = int(self.device)
        c.precision = {"f32": _abi.NDS_F32, "fp32": _abi.NDS_F32, "float32": _abi.NDS_F32,
                       "f64": _abi.NDS_F64, "fp64": _abi.NDS_F64, "float64": _abi.NDS_F64}[
            str(self.precision).lower()]
        c.ndim = self.ndim
        c.method = {"committor": _abi.NDS_METHOD_COMMITTOR, "mc": _abi.NDS_METHOD_MC,
                    "wl-mc": _abi.NDS_METHOD_WL}.get(self.method, _abi.NDS_METHOD_MD)
        if self.method in ("mc", "wl-mc"):
            for d in range(self.ndim):
                c.mc_bounds[d] = float(self.mc_bounds[d])
                c.mc_global_bounds[d] = float(self.mc_global_bounds[d])
            c.mc_propose_all = int(self.mc_propose_dim == "all")
            c.mc_periodic = int(self.mc_bound_cond == "periodic")
        c.integrator = _abi.INTEGRATORS[self.integrate]
        c.potential = _abi.POTENTIALS[self.potential][0]
        c.colvar = _abi.COLVARS[self.colvar][0]
        c.true_colvar = _abi.COLVARS[self.true_colvar][0]
        c.noise_mode = self.noise_mode
        c.steps_per_launch = int(self.steps_per_launch)
        c.n_walkers = int(self.n_walkers)
        c.walker_offset = int(self.walker_offset)
        c.n_walkers_global = int(self.n_walkers_global or self.n_walkers)
        c.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        c.temperature = float(self.temperature)
        c.initial_temperature = float(self.initial_temperature)
        c.mass = float(self.mass)
        c.dt = float(self.dt)
        c.gamma = float(self.gamma)
        c.rescale_freq = -1.0 if self.rescale_freq is None else float(self.rescale_freq)
        for i, v in enumerate(self.pot_params):
            c.pot_params[i] = float(v)
        c.n_umb = len(self.umb_k)
        c.umb_per_walker = int(self.umb_per_walker)
        for u in range(c.n_umb):
            for k in range(self.colvardim):
                c.umb_k[u][k] = float(self.umb_k[u][k])
                c.umb_r0[u][k] = float(self.umb_r0[u][k])
        c.mtd_enabled = int("mtd" in self.biases)
        c.mtd_shared = int(self.mtd_shared)
        if c.mtd_enabled:
            c.mtd_w = float(self.mtd_w)
            for k in range(self.colvardim):
                c.mtd_sigma[k] = float(self.mtd_sigma[k])
            c.mtd_dep_freq = float(self.mtd_dep_freq)
            c.mtd_biasf = float(self.mtd_biasf)
            c.mtd_capacity = int(self.mtd_capacity)
            c.mtd_adapsig_geo = int(self.mtd_adapsig_geo)
            c.mtd_adapsig_hist = int(self.mtd_adapsig_hist)
            c.mtd_grid = int(self.mtd_grid)
            if self.mtd_grid:
                for k in range(self.colvardim):
                    c.mtd_grid_lo[k] = float(self.mtd_grid_lo[k])
                    c.mtd_grid_hi[k] = float(self.mtd_grid_hi[k])
                    c.mtd_grid_n[k] = int(self.mtd_grid_n[k])
                if self.colvardim == 1:
                    c.mtd_grid_n[1] = 1
        c.pe_enabled = int("pe" in self.biases)
        c.pe_thred = float(self.pe_thred)
        c.n_basins = int(self.n_basins)
        c.diffusion_time = int(self.diffusion_time)
        if self.criteria is not None:
            for b in range(self.n_basins):
                for k in range(self.colvardim):
                    c.criteria[b][k][0] = float(self.criteria[b][k][0])
                    c.criteria[b][k][1] = float(self.criteria[b][k][1])
        c.emin = float(self.emin)
        c.min_gtol = float(self.min_gtol)
        c.min_maxiter = int(self.min_maxiter if self.min_maxiter is not None else self.steps * 10)
        if self.hist_n is not None:
            for k in range(min(self.colvardim, 2)):
                c.hist_n[k] = int(self.hist_n[k])
                c.hist_lo[k] = float(self.hist_lo[k])
                c.hist_hi[k] = float(self.hist_hi[k])
            if self.colvardim == 1:
                c.hist_n[1] = 1
        c.dump_freq = int(self.dump_freq) if (self.dump or self.hist_n is not None) else 0
        c.dump_walkers = int(min(self.dump_walkers, self.n_walkers)) if self.dump else 0
        c.dump_capacity = int(self.dump_capacity)
        c.stat_freq = int(self.stat_freq)
        return c


def colvar_dim(name: str, ndim: int) -> int:
    _, need_ndim, cd = _abi.COLVARS[name]
    if need_ndim is not None and need_ndim != ndim:
        raise ValueError(f"colvar {name} is defined for ndim={need_ndim}, not ndim={ndim}")
    return ndim if cd is None else cd


def parse(config: dict) -> RunSpec:
    """Reference kwargs (``NDRun(**config)``, ndrun.py:77-97) -> :class:`RunSpec`."""
    cfg = dict(config)
    if "ndim" not in cfg:
        raise TypeError("NDRun() missing 1 required positional argument: 'ndim'")
    ndim = int(cfg["ndim"])
    spec = RunSpec(ndim=ndim)
    spec.temperature = cfg.get("temperature", 300.0)
    spec.initial_temperature = cfg.get("initial_temperature", spec.temperature)  # ndrun.py:149
    spec.mass = cfg.get("mass", 5)
    spec.steps = int(cfg.get("steps", 100))
    spec.method = cfg.get("method", "md")
    seed = cfg.get("seed", None)
    spec.seed = int.from_bytes(os.urandom(4), "little") if seed is None else int(seed)
    spec.screen = bool(cfg.get("screen", True))
    spec.root = cfg.get("root", "./")
    spec.run_name = cfg.get("run_name", "ndrun")

    # ---- engine (engines/__init__.py:11-31)
    method = spec.method
    if method == "md":
        eng_prefix = ["md", "engine"]
    elif method == "committor":
        eng_prefix = ["engine_method", "committor", "engine"]
    elif method == "minimize":
        eng_prefix = ["minimize", "engine"]
    elif method in ("mc", "wl-mc"):
        eng_prefix = [method, "engine"]
    elif method in ("read_dump",):
        raise NotImplementedError(
            f"method '{method}' is not part of the B200 multi-walker MD path (SURVEY.md section 8f)")
    else:
        raise NameError(f"method type ``{method}'' undefined")  # engines/__init__.py:28
    spec.dt = pick(cfg, "dt", eng_prefix, 0.5)
    spec.gamma = pick(cfg, "gamma", eng_prefix, 0.005)
    spec.integrate = pick(cfg, "integrate", eng_prefix, "langevin")
    assert spec.integrate in ["langevin", "2nd-langevin", "rescale", "nve"]  # md.py:59
    if method == "minimize" or (method == "committor" and spec.temperature == 0):
        # engines/__init__.py:22-23: the committor's inner engine is Minimize at temperature 0
        spec.integrate = "minimize"
        spec.min_gtol = pick(cfg, "gtol", eng_prefix, 1.0e-5)
        spec.min_maxiter = pick(cfg, "maxiter", eng_prefix, None)
    spec.rescale_freq = pick(cfg, "rescale_freq", eng_prefix, None)
    if spec.integrate == "rescale" and spec.rescale_freq is None:
        raise TypeError("integrate: rescale needs rescale_freq (md.py:220)")

    if method in ("mc", "wl-mc"):  # engines/mc.py:7-32 (WangLandau subclasses MC)
        spec.mc_bounds = np.array(pick(cfg, "bounds", eng_prefix, [1.0, 1.0]), dtype=float).reshape(-1)
        spec.mc_global_bounds = np.array(pick(cfg, "global_bounds", eng_prefix, [20.0, 20.0]), dtype=float).reshape(-1)
        spec.mc_propose_dim = pick(cfg, "propose_dim", eng_prefix, "single")
        spec.mc_bound_cond = pick(cfg, "bound_cond", eng_prefix, "periodic")
        if len(spec.mc_bounds) != ndim:
            raise NameError("wrong input shape of mc_bounds")  # mc.py:31-32
        if spec.mc_bound_cond == "periodic" and len(spec.mc_global_bounds) != ndim:
            raise ValueError("global_bounds must have ndim entries (mc.py:60 broadcasts x %= global_bounds)")
        if spec.mc_propose_dim not in ("single", "all"):
            raise ValueError("propose_dim must be 'single' or 'all' (mc.py:51-55)")
        if len(spec.mc_global_bounds) != ndim:
            spec.mc_global_bounds = np.ones(ndim)

    # ---- potential (potentials/_build.py:4-21)
    pot = _builtin_name(cfg.get("potential", "Gaussian"), "potentials", _abi.POTENTIALS, "potential")
    spec.potential = pot
    _, pot_ndim, needs_tcv = _abi.POTENTIALS[pot]
    assert pot_ndim is None or ndim == pot_ndim, f"potential {pot} has ndim={pot_ndim} (potential.py:12)"
    pp = [0.0] * 16
    if pot.startswith("Mueller"):
        from .constant import lscale
        xc = pick(cfg, "x_center", "potential", np.array([1, 0, -0.5, -1]) * lscale + 32)
        yc = pick(cfg, "y_center", "potential", np.array([0, 0.5, 1.5, 1]) * lscale + 8)
        pp[0:4] = [float(v) for v in np.asarray(xc, dtype=float).reshape(-1)[:4]]
        pp[4:8] = [float(v) for v in np.asarray(yc, dtype=float).reshape(-1)[:4]]
    elif pot == "DoubleWell1d":
        pp[0] = pick(cfg, "K1", "potential", 1.0)
        pp[1] = pick(cfg, "K2", "potential", 1.0)
        pp[2] = pick(cfg, "x1", "potential", 10.0)
        pp[3] = pick(cfg, "x2", "potential", 30.0)
        pp[4] = pick(cfg, "sigma", "potential", 10)
    elif pot == "DoubleWell":
        pp[0] = pick(cfg, "K1", "potential", 1.0)
        pp[1] = pick(cfg, "K2", "potential", 1.0)
        pp[2:4] = list(np.asarray(pick(cfg, "x1", "potential", [20.0, 10.0]), dtype=float))
        pp[4:6] = list(np.asarray(pick(cfg, "x2", "potential", [20.0, 30.0]), dtype=float))
        pp[6] = pick(cfg, "sigma", "potential", 10)
    elif pot == "HarmonicQuartic":  # extension: V = sum_d k_d u_d^2 / 2 + q_d u_d^4 / 4, u = x - center
        def per_dim(key, default):
            v = np.asarray(pick(cfg, key, "potential", default), dtype=float).reshape(-1)
            if v.size == 1:
                v = np.repeat(v, ndim)
            if v.size != ndim:
                raise ValueError(f"potential_{key} needs 1 or ndim={ndim} values")
            return [float(t) for t in v]
        pp[0:ndim] = per_dim("k", 0.01)
        pp[5:5 + ndim] = per_dim("q", 0.0)
        pp[10:10 + ndim] = per_dim("center", 0.0)
    spec.pot_params = [float(v) for v in pp]

    # ---- colvars (colvars/_build.py:4-23)
    spec.colvar = _builtin_name(cfg.get("colvar_name", "Original"), "colvars", _abi.COLVARS, "colvar")
    spec.true_colvar = _builtin_name(cfg.get("true_colvar_name", "Original"), "colvars",
                                     _abi.COLVARS, "colvar")
    spec.colvardim = colvar_dim(spec.colvar, ndim)
    if spec.true_colvar in ("Path", "Path5dto2d"):
        raise NotImplementedError("path variables are supported as colvar_name, not as true_colvar_name")
    if spec.colvar in ("Path", "Path5dto2d"):  # colvars/path.py:32-60
        fname = pick(cfg, "file_name", "colvar", None)
        spec.path_sigma = float(pick(cfg, "sigma", "colvar", 1.0))
        ref = pick(cfg, "ref", "colvar", None)
        ind = pick(cfg, "ref_ind", "colvar", None)
        if fname is not None:
            data = np.loadtxt(fname)
            ref, ind = data[:, 1:], data[:, 0]
        if ref is None or ind is None:
            raise TypeError("path colvar needs colvar_file_name or colvar_ref + colvar_ref_ind")
        spec.path_ind = np.asarray(ind, dtype=float).reshape(-1)
        spec.path_ref = np.asarray(ref, dtype=float).reshape(len(spec.path_ind), -1)
        want = ndim if spec.colvar == "Path" else 2
        if spec.path_ref.shape[1] != want:
            raise ValueError(f"{spec.colvar}: reference points have {spec.path_ref.shape[1]} columns, expected {want}")
    tcd = colvar_dim(spec.true_colvar, ndim)
    if needs_tcv and tcd != 2:
        raise ValueError(f"potential {pot} is evaluated in a 2-D true_colvar space; "
                         f"true_colvar_name={spec.true_colvar} has colvardim {tcd}")

    # ---- x0 (ndrun.py:148,165-168)
    x0 = np.array(cfg.get("x0", [0.0, 0.0])).reshape((-1))
    if x0.dtype.kind in "US" or x0.shape[0] != ndim:
        raise NameError("wrong x0 dimension")
    spec.x0 = x0.astype(float)

    # ---- biases (bias/__init__.py:15-40)
    biases = list(cfg.get("biases", []) or [])
    spec.biases = biases
    cd = spec.colvardim
    for b in biases:
        if b == "mtd":  # bias/gaus.py:48-128
            adapsig, geo = pick(cfg, "adapsig", "mtd", False), pick(cfg, "adapsig_geo", "mtd", False)
            if adapsig and not geo:  # gaus.py:166-183,193-212: variance from the Stat history of the colvar
                if cd != 1:
                    raise ValueError("mtd_adapsig needs a 1-D colvar: for colvardim > 1 the reference sums "
                                     "exp(-dc_i dc_j / (2 S_ij)) over the ELEMENTS of the covariance matrix "
                                     "(gaus.py:268-277), which diverges for negative covariances")
                spec.mtd_adapsig_hist = True
            if adapsig and geo:      # gaus.py:162-165,189-192: every hill gets the variance J.J^T (* sigma^2)
                if cd != 1:
                    raise ValueError("mtd_adapsig_geo needs a 1-D colvar: for colvardim > 1 the reference takes the "
                                     "element-wise reciprocal of J.J^T (gaus.py:268-275), which is inf / NaN for "
                                     "the built-in colvars")
                spec.mtd_adapsig_geo = True
            spec.mtd_w = pick(cfg, "w", "mtd", None)
            sigma = pick(cfg, "sigma", "mtd", None)
            if isinstance(sigma, float):  # gaus.py:76-78
                sigma = np.array([sigma for _ in range(cd)])
            else:
                sigma = np.array(sigma, dtype=float)
                if sigma.ndim != 1 or sigma.shape[0] != cd:
                    raise NameError("wrong input shape of gaussian width sigma")  # gaus.py:82
            spec.mtd_sigma = sigma
            spec.mtd_dep_freq = pick(cfg, "dep_freq", "mtd", 10)
            spec.mtd_biasf = pick(cfg, "biasf", "mtd", 1.0)
            if not spec.mtd_biasf > 1.0:
                raise ValueError("mtd_biasf must be > 1: the reference divides by kB*(biasf-1)*T "
                                 "(bias/gaus.py:102-103,187)")
            if spec.mtd_w is None:
                raise TypeError("mtd_w is required")
        elif b == "umb":  # bias/umb.py:9-18
            umb_n = int(cfg.get("umb_n", 1))
            if umb_n > _abi.NDS_MAX_UMB:
                raise ValueError(f"umb_n > {_abi.NDS_MAX_UMB}")
            for i in range(umb_n):
                k = pick(cfg, "k", [f"umb_{i}", "umb"])
                r0 = pick(cfg, "r0", [f"umb_{i}", "umb"])
                if k is _MISSING or r0 is _MISSING:
                    raise TypeError("Harmonic.__init__() missing required arguments 'k' / 'r0'")
                r0 = np.array(r0, dtype=float).reshape(-1)
                assert r0.shape[0] == cd  # umb.py:18
                k = np.broadcast_to(np.array(k, dtype=float).reshape(-1), (cd,)).copy()
                spec.umb_k.append(k)
                spec.umb_r0.append(r0)
        elif b == "pe":  # bias/pe.py:13-17
            spec.pe_thred = pick(cfg, "thred", "pe", 0.0)
        else:
            raise UnboundLocalError(f"unknown bias {b}")  # bias/__init__.py:40 would fail on `instance`

    if method in ("mc", "wl-mc"):
        cfg["dump"] = False  # the reference dumps only on accepted moves (ndrun.py:234-251): not recorded

    # ---- committor (engines/committor.py:31-75)
    if method == "committor":
        spec.n_basins = int(pick(cfg, "n_basins", ["committor", "engine"]))
        spec.criteria = np.array(pick(cfg, "criteria", ["committor", "engine"]), dtype=float)
        spec.diffusion_time = int(pick(cfg, "diffusion_time", ["committor", "engine"], 0))
        spec.emin = float(pick(cfg, "emin", ["committor", "engine"], -1.2))
        assert len(spec.criteria) == spec.n_basins  # committor.py:57
        for crit in spec.criteria:
            assert len(crit) == cd  # committor.py:73
        if spec.n_basins > _abi.NDS_MAX_BASINS:
            raise ValueError(f"n_basins > {_abi.NDS_MAX_BASINS}")

    # ---- dumps / stat (io/dump.py:34-39, io/stat.py:7-12)
    spec.dump = bool(cfg.get("dump", True))
    spec.dump_freq = int(pick(cfg, "freq", "dump", 1))
    spec.stat_freq = int(pick(cfg, "freq", "stat", 1))

    # ---- new keys of the multi-walker engine
    spec.n_walkers = int(cfg.get("n_walkers", 1))
    spec.walker_offset = int(cfg.get("walker_offset", 0))
    spec.n_walkers_global = int(cfg.get("n_walkers_global", spec.n_walkers))
    spec.precision = cfg.get("precision", "f64")
    spec.steps_per_launch = int(cfg.get("steps_per_launch", 1000))
    spec.device = int(cfg.get("device", 0))
    spec.noise_mode = int(cfg.get("noise_mode", _abi.NDS_NOISE_PHILOX))
    spec.dump_walkers = int(cfg.get("dump_walkers", 1))
    # under MC NDRun.time advances on accepted moves only (ndrun.py:234-237): every walker deposits on its own
    # clock, so the tables are per walker
    spec.mtd_shared = bool(cfg.get("mtd_shared", method not in ("mc", "wl-mc")))
    if "mtd" in biases and method == "mc" and spec.mtd_shared:
        raise ValueError("metadynamics under method: mc needs mtd_shared: false (per-walker deposition clocks)")
    spec.umb_per_walker = bool(cfg.get("umb_per_walker", False))
    if "mtd" in biases and cfg.get("mtd_grid", False):
        # grid-accumulated bias (new keys): box [mtd_grid_lo, mtd_grid_hi] per colvar -- default: the reference's
        # `boundary` / `plot_boundary` key (plot.py:42-46) -- and spacing mtd_grid_spacing (default sigma / 4)
        if spec.mtd_adapsig_geo or spec.mtd_adapsig_hist or not spec.mtd_shared or cd > 2:
            raise ValueError("mtd_grid needs a shared table, a fixed sigma and colvardim <= 2")
        box = cfg.get("mtd_grid_box", cfg.get("plot_boundary", cfg.get("boundary", None)))
        lo, hi = cfg.get("mtd_grid_lo", None), cfg.get("mtd_grid_hi", None)
        if lo is None or hi is None:
            if box is None:
                raise TypeError("mtd_grid needs mtd_grid_lo / mtd_grid_hi (or boundary)")
            box = np.asarray(box, dtype=float).reshape(-1, 2)
            lo, hi = box[:cd, 0], box[:cd, 1]
        spec.mtd_grid_lo = np.asarray(lo, dtype=float).reshape(-1)[:cd]
        spec.mtd_grid_hi = np.asarray(hi, dtype=float).reshape(-1)[:cd]
        if len(spec.mtd_grid_lo) != cd or len(spec.mtd_grid_hi) != cd:
            raise ValueError("mtd_grid_lo / mtd_grid_hi need colvardim entries")
        n = cfg.get("mtd_grid_n", None)
        if n is None:
            h = np.broadcast_to(np.asarray(cfg.get("mtd_grid_spacing", spec.mtd_sigma / 4.0), dtype=float), (cd,))
            n = np.ceil((spec.mtd_grid_hi - spec.mtd_grid_lo) / h).astype(int) + 1
        spec.mtd_grid_n = np.broadcast_to(np.asarray(n, dtype=int), (cd,)).copy()
        spec.mtd_grid = True
    if "mtd" in biases:
        cap = cfg.get("mtd_capacity", None)
        if cap is None:
            n_dep = int(np.floor(spec.steps * spec.dt / spec.mtd_dep_freq)) + 2
            per = spec.n_walkers_global if spec.mtd_shared else 1
            cap = n_dep * per
        spec.mtd_capacity = int(cap)
    if cfg.get("hist_n", None) is not None:
        # colvar histogram accumulated inside the step kernel at dump_freq cadence over ALL walkers (new keys); the
        # box defaults to the reference's plot boundary (plot.py:42-46, 562-572)
        if cd > 2:
            raise ValueError("hist_n: colvardim > 2")
        box = cfg.get("plot_boundary", cfg.get("boundary", None))
        lo, hi = cfg.get("hist_lo", None), cfg.get("hist_hi", None)
        if lo is None or hi is None:
            if box is None:
                raise TypeError("hist_n needs hist_lo / hist_hi (or boundary)")
            box = np.asarray(box, dtype=float).reshape(-1, 2)
            lo, hi = box[:cd, 0], box[:cd, 1]
        spec.hist_lo = np.asarray(lo, dtype=float).reshape(-1)[:cd]
        spec.hist_hi = np.asarray(hi, dtype=float).reshape(-1)[:cd]
        spec.hist_n = np.broadcast_to(np.asarray(cfg["hist_n"], dtype=int), (cd,)).copy()
    if spec.dump:
        # dump rows + the lagged stat sample before each of them
        spec.dump_capacity = int(cfg.get("dump_capacity", 2 * (spec.steps // max(spec.dump_freq, 1)) + 2))
    return spec
