"""TEST INFRASTRUCTURE ONLY: ctypes wrapper around ``oracle/libnds_oracle.so`` (``nds_oracle.c``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module, and only as the checker or the timed CPU baseline.  The product
package ``ndsimulator_b200`` never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ndsimulator_b200._abi import NdsConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnds_oracle.so")
_lib = None
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "nds_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ndsb200.h")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnds_oracle.so"], env=env,
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(NdsConfig), C.POINTER(C.c_uint32)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_constants.argtypes = [C.c_void_p, _dp]
        L.oracle_set_state.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_set_walker_bias.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_set_path.argtypes = [C.c_void_p, C.c_int64, C.c_int, _dp, _dp, C.c_double]
        L.oracle_set_noise.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64]
        L.oracle_begin.argtypes = [C.c_void_p]
        L.oracle_frame_width.argtypes = [C.c_void_p]
        L.oracle_frame_width.restype = C.c_int
        L.oracle_run.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int64]
        L.oracle_run.restype = C.c_int64
        L.oracle_step.argtypes = [C.c_void_p]
        L.oracle_step.restype = C.c_int64
        L.oracle_time.argtypes = [C.c_void_p]
        L.oracle_time.restype = C.c_double
        L.oracle_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]
        L.oracle_get_commit.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.oracle_get_vr.argtypes = [C.c_void_p, _dp]
        L.oracle_get_mc.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.oracle_get_wl_histogram.argtypes = [C.c_void_p, C.c_int64, _dp]
        L.oracle_n_hills.argtypes = [C.c_void_p, C.c_int64]
        L.oracle_n_hills.restype = C.c_int64
        L.oracle_get_hills.argtypes = [C.c_void_p, C.c_int64, _dp, _dp]
        L.oracle_set_hills.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _dp, _dp]
        L.oracle_get_hill_sigma2.argtypes = [C.c_void_p, C.c_int64, _dp]
        L.oracle_set_hill_sigma2.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _dp]
        L.oracle_eval_at.argtypes = [C.c_void_p, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp]
        L.oracle_fes_grid.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int64,
                                      C.c_double, C.c_double, C.c_int64, _dp]
        L.oracle_rng_normals.argtypes = [C.c_uint32, C.c_int64, _dp]
        L.oracle_rng_uniforms.argtypes = [C.c_uint32, C.c_int64, _dp]
        L.oracle_map50to2d.argtypes = [_dp, C.c_int64, _dp, _dp]
        L.oracle_colvar_dim.argtypes = [C.c_int, C.c_int]
        L.oracle_colvar_dim.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class OracleSim:
    """M independent (or hill-sharing) walkers stepped by the C restatement of the reference."""

    FIELDS = ("x", "v", "f", "fixf", "colv", "pe", "ke", "biase", "totale", "T", "vr")

    def __init__(self, cfg: NdsConfig, seeds=None):
        self.L = lib()
        self.cfg = cfg
        self.M = int(cfg.n_walkers)
        self.ndim = int(cfg.ndim)
        self.cd = self.L.oracle_colvar_dim(cfg.colvar, cfg.ndim)
        sp = None
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
            assert self._seeds.shape == (self.M,)
            sp = self._seeds.ctypes.data_as(C.POINTER(C.c_uint32))
        self.h = C.c_void_p(self.L.oracle_create(C.byref(cfg), sp))
        self._noise = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_destroy(self.h)
            self.h = None

    def set_threads(self, n):
        self.L.oracle_set_threads(self.h, int(n))

    def constants(self):
        out = np.zeros(8)
        self.L.oracle_constants(self.h, _p(out))
        return dict(zip(("m", "kBT", "c1", "c2", "c3", "c4", "c5", "v_target"), out))

    def set_state(self, x, v=None):
        x = _f64(np.asarray(x).reshape(self.ndim, self.M))
        v = None if v is None else _f64(np.asarray(v).reshape(self.ndim, self.M))
        self.L.oracle_set_state(self.h, _p(x), _p(v))

    def set_walker_bias(self, k, r0):
        k = _f64(np.asarray(k).reshape(self.cd, self.M))
        r0 = _f64(np.asarray(r0).reshape(self.cd, self.M))
        self.L.oracle_set_walker_bias(self.h, _p(k), _p(r0))

    def set_path(self, ref, ref_ind, sigma=1.0):
        ref = _f64(np.asarray(ref, dtype=float).reshape(len(ref_ind), -1))
        ind = _f64(np.asarray(ref_ind, dtype=float).reshape(-1))
        self.L.oracle_set_path(self.h, len(ind), ref.shape[1], _p(ref), _p(ind), float(sigma))

    def set_noise(self, noise, first_step=0):
        self._noise = _f64(noise)  # [nsteps][nn][M], kept alive
        self.L.oracle_set_noise(self.h, _p(self._noise), int(first_step), int(self._noise.shape[0]))

    def begin(self):
        self.L.oracle_begin(self.h)

    def run(self, nsteps, traj_every=0):
        """Returns the trajectory ``[nframes][width][M]`` when ``traj_every`` > 0."""
        traj = None
        if traj_every:
            W = self.L.oracle_frame_width(self.h)
            traj = np.zeros((nsteps // traj_every, W, self.M))
        self.L.oracle_run(self.h, int(nsteps), _p(traj), int(traj_every))
        return traj

    def split_frame(self, frames):
        """dict of named views into a trajectory array returned by :meth:`run`."""
        nd, cd = self.ndim, self.cd
        o, out = 0, {}
        for name, n in (("x", nd), ("v", nd), ("f", nd), ("fixf", nd), ("colv", cd)):
            out[name] = frames[:, o:o + n]
            o += n
        for name in ("pe", "ke", "biase", "totale", "T", "vr"):
            out[name] = frames[:, o]
            o += 1
        return out

    @property
    def step(self):
        return self.L.oracle_step(self.h)

    @property
    def time(self):
        return self.L.oracle_time(self.h)

    def get_state(self):
        M, nd, cd = self.M, self.ndim, self.cd
        x, v, f, fixf = (np.zeros((nd, M)) for _ in range(4))
        colv, thermo = np.zeros((cd, M)), np.zeros((5, M))
        self.L.oracle_get_state(self.h, _p(x), _p(v), _p(f), _p(fixf), _p(colv), _p(thermo))
        return dict(x=x, v=v, f=f, fixf=fixf, colv=colv, T=thermo[0], pe=thermo[1], ke=thermo[2],
                    biase=thermo[3], totale=thermo[4])

    def get_commit(self):
        basin = np.zeros(self.M, dtype=np.int32)
        exit_step = np.zeros(self.M, dtype=np.int64)
        self.L.oracle_get_commit(self.h, basin.ctypes.data_as(C.POINTER(C.c_int32)),
                                 exit_step.ctypes.data_as(C.POINTER(C.c_int64)))
        return basin, exit_step

    def get_mc(self):
        acc = np.zeros(self.M, dtype=np.int64)
        self.L.oracle_get_mc(self.h, acc.ctypes.data_as(C.POINTER(C.c_int64)))
        return acc

    def get_wl_histogram(self, walker=0):
        h = np.zeros(101)
        self.L.oracle_get_wl_histogram(self.h, int(walker), _p(h))
        return h

    def get_vr(self):
        vr = np.zeros(self.M)
        self.L.oracle_get_vr(self.h, _p(vr))
        return vr

    def get_hills(self, walker=0):
        n = self.L.oracle_n_hills(self.h, walker)
        c, w = np.zeros((n, self.cd)), np.zeros(n)
        self.L.oracle_get_hills(self.h, walker, _p(c), _p(w))
        return c, w

    def set_hills(self, centers, heights, walker=0):
        c = _f64(np.asarray(centers).reshape(-1, self.cd))
        w = _f64(heights)
        self.L.oracle_set_hills(self.h, walker, len(w), _p(c), _p(w))

    def get_hill_sigma2(self, walker=0):
        n = self.L.oracle_n_hills(self.h, walker)
        s2 = np.zeros(n)
        self.L.oracle_get_hill_sigma2(self.h, walker, _p(s2))
        return s2

    def set_hill_sigma2(self, sigma2, walker=0):
        s2 = _f64(sigma2)
        self.L.oracle_set_hill_sigma2(self.h, walker, len(s2), _p(s2))

    def eval_at(self, x):
        x = _f64(np.asarray(x).reshape(self.ndim, -1))
        n, nd, cd = x.shape[1], self.ndim, self.cd
        V, Vb = np.zeros(n), np.zeros(n)
        F, Fb = np.zeros((nd, n)), np.zeros((nd, n))
        colv, J = np.zeros((cd, n)), np.zeros((cd, nd, n))
        self.L.oracle_eval_at(self.h, _p(x), n, _p(V), _p(F), _p(colv), _p(J), _p(Vb), _p(Fb))
        return dict(V=V, F=F, colv=colv, J=J, Vb=Vb, Fb=Fb)

    def fes_grid(self, x0, dx, nx, y0=0.0, dy=1.0, ny=1, walker=0):
        g = np.zeros((ny, nx))
        self.L.oracle_fes_grid(self.h, walker, x0, dx, nx, y0, dy, ny, _p(g))
        return g


def rng_normals(seed, n):
    out = np.zeros(n)
    lib().oracle_rng_normals(int(seed), n, _p(out))
    return out


def rng_uniforms(seed, n):
    out = np.zeros(n)
    lib().oracle_rng_uniforms(int(seed), n, _p(out))
    return out


def max_threads():
    return lib().oracle_max_threads()


def map50to2d(x):
    """Map50to2d (colvars/misc.py:19-156) at x[50][n] through the dense C restatement -> dict(c=[2][n], J=[2][50][n])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[1]
    c, J = np.zeros((2, n)), np.zeros((2, 50, n))
    lib().oracle_map50to2d(_p(x), n, _p(c), _p(J))
    return dict(c=c, J=J)
