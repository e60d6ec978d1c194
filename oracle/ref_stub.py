"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-ins that let the UNMODIFIED Python reference (``/root/reference/ndsimulator``) be imported
in the authoring container, where its two non-arithmetic dependencies are absent:

* ``pyfile_utils`` (PyPI ``pyfile-utils>=1.0.5``, reference ``setup.py:24``; not vendored, no pin) --
  config-key -> constructor mapping, dotted-name import, log-file creation.  No arithmetic.
  Call sites: ``ndrun.py:40,105-110``, ``engines/__init__.py:30``, ``engines/committor.py:50``,
  ``bias/__init__.py:28-36``, ``potentials/_build.py:13-20``, ``colvars/_build.py:13-22``,
  ``io/dump.py:54-69``, ``scripts/run.py:74``.
* ``matplotlib`` -- mocked; callers must set ``sim.plot = None`` after construction because
  ``NDRun`` unconditionally builds a ``Plot`` (``ndrun.py:140``).

The semantics of ``instantiate`` restated here (fill constructor parameters from bare keys, then from
``<prefix>_<key>``, later prefixes winning; unknown keys ignored; a class-valued parameter with a
sibling ``<param>_kwargs`` parameter collects that class's kwargs recursively) are recalled from the
library's documented behaviour; they cannot be verified against its source in this container.

Used only by ``oracle/make_golden.py`` (fixture generation) and by the optional cross-checks in
``tests/`` that are skipped when ``/root/reference`` does not exist (e.g. on the GPU box).
"""
import importlib
import inspect
import logging
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("NDS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ndsimulator"))


def _load_callable(name):
    if callable(name):
        return name
    if not isinstance(name, str) or "." not in name:
        raise ValueError(f"cannot load {name!r}")
    module_name, attr = name.rsplit(".", 1)
    module = importlib.import_module(module_name)
    return getattr(module, attr)


def _instantiate(builder, prefix=None, positional_args=None, optional_args=None,
                 all_args=None, remove_kwargs=True, return_args_only=False):
    if prefix is None:
        prefixes = []
    elif isinstance(prefix, str):
        prefixes = [prefix]
    else:
        prefixes = list(prefix)
    optional_args = {} if optional_args is None else optional_args
    positional_args = {} if positional_args is None else positional_args
    sig = inspect.signature(builder.__init__ if inspect.isclass(builder) else builder)
    names = [k for k in sig.parameters if k not in ("self", "kwargs", "args")]
    kwargs = {}
    for src in (all_args or {}, optional_args):
        for k in names:
            if k in src:
                kwargs[k] = src[k]
        for p in prefixes:
            for k in names:
                if f"{p}_{k}" in src:
                    kwargs[k] = src[f"{p}_{k}"]
    kwargs.update({k: v for k, v in positional_args.items() if k in names})
    # class-valued parameter + "<param>_kwargs" sibling: collect the inner class's kwargs
    for k in list(kwargs):
        if inspect.isclass(kwargs[k]) and f"{k}_kwargs" in names:
            inner = dict(kwargs.get(f"{k}_kwargs", {}) or {})
            _, inner_kwargs = _instantiate(kwargs[k], prefix=[k] + prefixes,
                                           optional_args={**optional_args, **inner},
                                           return_args_only=True)
            kwargs[f"{k}_kwargs"] = inner_kwargs
    if return_args_only:
        return None, kwargs
    return builder(**kwargs), kwargs


class _Output:
    def __init__(self, root="./", run_name="ndrun", **kw):
        self.root = root
        self.run_name = run_name
        self.workdir = os.path.join(root, run_name)
        os.makedirs(self.workdir, exist_ok=True)

    @classmethod
    def get_output(cls, kwargs):
        return cls(root=kwargs.get("root", "./"), run_name=kwargs.get("run_name", "ndrun"))

    def open_logfile(self, file_name, screen=False, propagate=False):
        name = f"{self.workdir}/{file_name}"
        logger = logging.getLogger(name)
        logger.propagate = propagate
        logger.setLevel(logging.INFO)
        for h in list(logger.handlers):
            logger.removeHandler(h)
        fh = logging.FileHandler(name, mode="w")
        fh.setFormatter(logging.Formatter("%(message)s"))
        logger.addHandler(fh)
        return name


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @classmethod
    def from_file(cls, path, defaults=None):
        import yaml
        with open(path) as f:
            d = yaml.safe_load(f)
        out = cls(dict(defaults or {}))
        out.update(d)
        return out


def install():
    """Register the stand-ins and put the reference on ``sys.path``.  Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "pyfile_utils" not in sys.modules:
        m = types.ModuleType("pyfile_utils")
        m.instantiate = _instantiate
        m.load_callable = _load_callable
        m.Output = _Output
        m.Config = _Config
        sys.modules["pyfile_utils"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.collections",
                 "matplotlib.cm", "matplotlib.animation"):
        sys.modules.setdefault(name, mock.MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def make_run(cfg: dict, v0=None):
    """Build ``NDRun(**cfg)`` from the shipped reference with plotting disabled; optionally inject v0."""
    install()
    import numpy as np
    from ndsimulator.ndrun import NDRun
    cfg = dict(cfg)
    cfg.setdefault("plot_boundary", [[0.0, 48.0], [0.0, 40.0]])
    cfg.setdefault("screen", False)
    sim = NDRun(**cfg)
    sim.plot = None
    if v0 is not None:
        sim.modify.set_velocity(v=np.array(v0, dtype=float))
    return sim
