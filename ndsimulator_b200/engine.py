"""Thin numpy-facing wrapper over the C ABI (``include/ndsb200.h``).  One :class:`Engine` = one
``nds_engine`` handle = M walkers on one GPU.  All arrays are float64, SoA ``[d][M]``."""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import NdsConfig

_dp = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class NdsError(RuntimeError):
    pass


class Engine:
    def __init__(self, cfg: NdsConfig):
        self.lib = _abi.load_library()
        self.cfg = cfg
        self.M = int(cfg.n_walkers)
        self.ndim = int(cfg.ndim)
        self.cd = self.lib.nds_colvar_dim(cfg.colvar, cfg.ndim)
        self.h = C.c_void_p()
        if self.lib.nds_create(C.byref(cfg), C.byref(self.h)):
            raise NdsError(self.lib.nds_last_error(None).decode())
        self._exchange_cb = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.nds_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise NdsError(self.lib.nds_last_error(self.h).decode())

    # ---- state
    def set_state(self, x, v=None):
        x = _f64(x, (self.ndim, self.M))
        v = None if v is None else _f64(v, (self.ndim, self.M))
        self._check(self.lib.nds_set_state(self.h, _p(x), _p(v)))

    def set_state_f32(self, x, v=None):
        """Same as :meth:`set_state` with float32 host arrays (direct copies for fp32 engines)."""
        fp = C.POINTER(C.c_float)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(self.ndim, self.M)
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float32).reshape(self.ndim, self.M)
        self._check(self.lib.nds_set_state_f32(self.h, x.ctypes.data_as(fp),
                                               None if v is None else v.ctypes.data_as(fp)))

    def get_state_f32(self, fields=("x", "v", "thermo")):
        fp = C.POINTER(C.c_float)
        M, nd, cd = self.M, self.ndim, self.cd
        shapes = dict(x=(nd, M), v=(nd, M), f=(nd, M), fixf=(nd, M), colv=(cd, M), thermo=(5, M))
        bufs = {k: np.empty(shapes[k], dtype=np.float32) for k in fields}
        args = [(bufs[k].ctypes.data_as(fp) if k in bufs else None) for k in ("x", "v", "f", "fixf", "colv", "thermo")]
        self._check(self.lib.nds_get_state_f32(self.h, *args))
        out = {k: bufs[k] for k in fields if k != "thermo"}
        if "thermo" in bufs:
            t = bufs["thermo"]
            out.update(T=t[0], pe=t[1], ke=t[2], biase=t[3], totale=t[4])
        return out

    def set_walker_bias(self, k, r0):
        k = _f64(k, (self.cd, self.M))
        r0 = _f64(r0, (self.cd, self.M))
        self._check(self.lib.nds_set_walker_bias(self.h, _p(k), _p(r0)))

    def set_path(self, ref, ref_ind, sigma=1.0):
        """Reference points of a Path / Path5dto2d colvar (colvars/path.py:32-71)."""
        ind = _f64(np.asarray(ref_ind, dtype=float).reshape(-1))
        ref = _f64(np.asarray(ref, dtype=float).reshape(len(ind), -1))
        self._check(self.lib.nds_set_path(self.h, len(ind), ref.shape[1], _p(ref), _p(ind), float(sigma)))

    def set_noise(self, noise, first_step=0):
        noise = _f64(noise)
        self._check(self.lib.nds_set_noise(self.h, _p(noise), int(first_step), int(noise.shape[0])))

    def begin(self):
        self._check(self.lib.nds_begin(self.h))

    def run(self, nsteps, sync=True):
        self._check(self.lib.nds_run(self.h, int(nsteps)))
        if sync:
            self.sync()

    def sync(self):
        self._check(self.lib.nds_sync(self.h))

    def get_state(self, fields=("x", "v", "f", "fixf", "colv", "thermo")):
        M, nd, cd = self.M, self.ndim, self.cd
        bufs = dict(x=np.empty((nd, M)), v=np.empty((nd, M)), f=np.empty((nd, M)),
                    fixf=np.empty((nd, M)), colv=np.empty((cd, M)), thermo=np.empty((5, M)))
        args = [(_p(bufs[k]) if k in fields else None) for k in ("x", "v", "f", "fixf", "colv", "thermo")]
        self._check(self.lib.nds_get_state(self.h, *args))
        out = {k: bufs[k] for k in fields if k != "thermo"}
        if "thermo" in fields:
            t = bufs["thermo"]
            out.update(T=t[0], pe=t[1], ke=t[2], biase=t[3], totale=t[4])
        return out

    @property
    def step(self):
        return int(self.lib.nds_step(self.h))

    @property
    def time(self):
        return float(self.lib.nds_time(self.h))

    # ---- deterministic evaluation
    def eval_at(self, x):
        x = _f64(x).reshape(self.ndim, -1)
        n, nd, cd = x.shape[1], self.ndim, self.cd
        V, Vb = np.empty(n), np.empty(n)
        F, Fb = np.empty((nd, n)), np.empty((nd, n))
        colv, J = np.empty((cd, n)), np.empty((cd, nd, n))
        self._check(self.lib.nds_eval_at(self.h, _p(x), n, _p(V), _p(F), _p(colv), _p(J), _p(Vb), _p(Fb)))
        return dict(V=V, F=F, colv=colv, J=J, Vb=Vb, Fb=Fb)

    # ---- metadynamics
    def n_hills(self, walker=0):
        return int(self.lib.nds_n_hills(self.h, walker))

    def get_hills(self, walker=0):
        n = self.n_hills(walker)
        c, w = np.empty((n, self.cd)), np.empty(n)
        nn = C.c_int64()
        self._check(self.lib.nds_get_hills(self.h, walker, C.byref(nn), _p(c), _p(w), n))
        return c, w

    def get_hill_sigma2(self, walker=0):
        """Per-hill variances (Gaussian._sigma_history for mtd_adapsig_geo; sigma^2 otherwise)."""
        n = self.n_hills(walker)
        s2 = np.empty(n)
        self._check(self.lib.nds_get_hill_sigma2(self.h, walker, _p(s2), n))
        return s2

    def set_hills(self, centers, heights, walker=-1, sigma2=None):
        c = _f64(centers).reshape(-1, self.cd)
        w = _f64(heights).reshape(-1)
        self._check(self.lib.nds_set_hills(self.h, walker, len(w), _p(c), _p(w)))
        if sigma2 is not None:
            s2 = _f64(sigma2).reshape(-1)
            self._check(self.lib.nds_set_hill_sigma2(self.h, walker, len(s2), _p(s2)))

    def get_vr(self):
        vr = np.empty(self.M)
        self._check(self.lib.nds_get_vr(self.h, _p(vr)))
        return vr

    def set_exchange(self, fn):
        """fn(dev_segment_ptr:int, bytes_per_rank:int, rank_offset_bytes:int) -> int (0 = ok)."""
        def tramp(_user, ptr, nbytes, off):
            try:
                return int(fn(ptr, nbytes, off) or 0)
            except Exception as e:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._exchange_cb = _abi.EXCHANGE_FN(tramp)
        self._check(self.lib.nds_set_exchange(self.h, self._exchange_cb, None))

    def hills_ipc_handle(self) -> bytes:
        """cudaIpcMemHandle_t (64 bytes) of this engine's hill table."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.nds_hills_ipc_handle(self.h, buf))
        return buf.raw

    def open_peer_tables(self, handles, rank):
        """handles: list of 64-byte IPC handles of every rank's table (same node), indexed by rank."""
        blob = b"".join(handles)
        self._check(self.lib.nds_open_peer_tables(self.h, blob, len(handles), int(rank)))

    def set_peer_tables(self, pointers, rank):
        """raw device pointers of every rank's table (engines living in this process)."""
        arr = (C.c_void_p * len(pointers))(*pointers)
        self._check(self.lib.nds_set_peer_tables(self.h, arr, len(pointers), int(rank)))

    # ---- native exchange / grid-accumulated bias
    def set_comm(self, unique_id: bytes = None, comm_ptr: int = None, rank=0, nranks=1):
        """Native NCCL exchange on the engine's stream (``nds_set_comm``): pass the 128-byte unique id of
        :func:`nccl_unique_id` (same on every rank) or a raw ``ncclComm_t``."""
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._check(self.lib.nds_set_comm(self.h, buf, C.c_void_p(comm_ptr) if comm_ptr else None, int(rank), int(nranks)))

    def set_peer_barrier(self, on_device=True):
        self._check(self.lib.nds_set_peer_barrier(self.h, int(bool(on_device))))

    def exchange_info(self):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.nds_exchange_info(self.h, C.byref(ms), C.byref(n)))
        return dict(exchange_ms=ms.value, n_exchanges=n.value)

    def table_checksum(self, first=0, n=None):
        n = self.n_hills() - first if n is None else n
        out = C.c_uint64()
        self._check(self.lib.nds_table_checksum(self.h, int(first), int(n), C.byref(out)))
        return int(out.value)

    def grid_shape(self):
        return (int(self.cfg.mtd_grid_n[1]) if self.cd > 1 else 1, int(self.cfg.mtd_grid_n[0]), 4)

    def get_mtd_grid(self):
        """Accumulated bias grid ``[ny][nx][4]`` = (V, h0 dV/dc0, h1 dV/dc1, h0 h1 d2V/dc0dc1)."""
        g = np.empty(self.grid_shape())
        self._check(self.lib.nds_get_mtd_grid(self.h, _p(g), g.size))
        return g

    def set_mtd_grid(self, g):
        g = _f64(g, self.grid_shape())
        self._check(self.lib.nds_set_mtd_grid(self.h, _p(g), g.size))

    def get_cv_history(self):
        """-> (ring [window][M], n_samples): the Stat.colv window the history-based adaptive sigma reads."""
        n, win = C.c_int64(), C.c_int64()
        self._check(self.lib.nds_get_cv_history(self.h, None, 0, C.byref(n), C.byref(win)))
        ring = np.empty((win.value, self.M))
        self._check(self.lib.nds_get_cv_history(self.h, _p(ring), ring.size, C.byref(n), C.byref(win)))
        return ring, int(n.value)

    def set_cv_history(self, ring, n_samples):
        ring = _f64(ring)
        self._check(self.lib.nds_set_cv_history(self.h, _p(ring), int(n_samples)))

    def mtd_grid_outside(self):
        return int(self.lib.nds_mtd_grid_outside(self.h))

    def submit(self, host_in, nsteps, host_out, chunks=1):
        """Asynchronous pass: ``host_in`` / ``host_out`` are raw addresses (ints) of pinned blocks in the engine's
        precision, ``[2 ndim][M]`` and ``[2 ndim + 5][M]``; complete with :meth:`wait`."""
        self._check(self.lib.nds_submit(self.h, C.c_void_p(host_in) if host_in else None, int(nsteps),
                                        C.c_void_p(host_out) if host_out else None, int(chunks)))

    def wait(self):
        self._check(self.lib.nds_wait(self.h))

    def fes_grid(self, x0, dx, nx, y0=0.0, dy=1.0, ny=1, walker=0):
        g = np.empty((ny, nx))
        self._check(self.lib.nds_fes_grid(self.h, walker, x0, dx, nx, y0, dy, ny, _p(g)))
        return g

    def colvar_histogram(self, x0, dx, nx, y0=0.0, dy=1.0, ny=1):
        counts = np.zeros((ny, nx), dtype=np.int64)
        self._check(self.lib.nds_colvar_histogram(self.h, x0, dx, nx, y0, dy, ny,
                                                  counts.ctypes.data_as(C.POINTER(C.c_int64))))
        return counts

    # ---- checkpoint / resume
    def get_clock(self):
        s, d, r = C.c_int64(), C.c_int64(), C.c_int64()
        t = C.c_double()
        self._check(self.lib.nds_get_clock(self.h, C.byref(s), C.byref(t), C.byref(d), C.byref(r)))
        return dict(step=s.value, time=t.value, dep_count=d.value, rescale_count=r.value)

    def set_clock(self, step, time, dep_count=0, rescale_count=0):
        self._check(self.lib.nds_set_clock(self.h, int(step), float(time), int(dep_count), int(rescale_count)))

    def set_vr(self, vr):
        vr = _f64(vr, (self.M,))
        self._check(self.lib.nds_set_vr(self.h, _p(vr)))

    def checkpoint(self):
        """Everything needed to continue the run bit for bit on a fresh engine of the same config."""
        st = self.get_state(fields=("x", "v", "thermo"))
        ck = dict(x=st["x"], v=st["v"], **self.get_clock())
        if self.cfg.method in (_abi.NDS_METHOD_MC, _abi.NDS_METHOD_WL):
            # Metropolis / Wang-Landau: accepted moves, ke / totale of the last event (they are not functions of (x, v):
            # mc.py:93-95 zeroes ke on the first accepted move), per-walker clocks + tables, histograms
            i64 = C.POINTER(C.c_int64)
            acc = np.zeros(self.M, dtype=np.int64)
            mtd, wl = bool(self.cfg.mtd_enabled), self.cfg.method == _abi.NDS_METHOD_WL
            tm = np.zeros(self.M) if mtd else None
            dep = np.zeros(self.M, dtype=np.int64) if mtd else None
            hist = np.zeros((101, self.M)) if wl else None
            mx = C.c_int64()
            self._check(self.lib.nds_get_mc_state(self.h, acc.ctypes.data_as(i64), _p(tm),
                                                  None if dep is None else dep.ctypes.data_as(i64), _p(hist), C.byref(mx)))
            ck.update(mc_accepted=acc, mc_thermo=np.stack([st[k] for k in ("T", "pe", "ke", "biase", "totale")]))
            if wl:
                ck["mc_wl"] = hist
            if mtd:
                n = int(mx.value)
                raw = np.empty(int(self.lib.nds_hill_table_size(self.h, n)))
                self._check(self.lib.nds_get_hill_table_raw(self.h, _p(raw), raw.size))
                ck.update(mc_time=tm, mc_dep=dep, hills_raw=raw, hills_n=n, vr=self.get_vr())
            return ck
        if self.cfg.method == _abi.NDS_METHOD_COMMITTOR:
            basin, exit_step = self.get_commit()
            alive = np.empty(self.M, dtype=np.uint8)
            self._check(self.lib.nds_get_alive(self.h, alive.ctypes.data_as(C.POINTER(C.c_uint8))))
            ck.update(basin=basin, exit_step=exit_step, alive=alive)
        if self.cfg.mtd_enabled:
            if self.cfg.mtd_shared:
                c, h = self.get_hills(0)
                ck.update(hills_centers=c, hills_heights=h)
                if self.cfg.mtd_adapsig_geo or self.cfg.mtd_adapsig_hist:
                    ck["hills_sigma2"] = self.get_hill_sigma2(0)
                if self.cfg.mtd_grid:  # the grid holds every rank's hills; the table only this rank's
                    ck["mtd_grid"] = self.get_mtd_grid()
            else:  # per-walker tables: the raw storage (engine layout)
                n = self.n_hills(0)
                raw = np.empty(int(self.lib.nds_hill_table_size(self.h, n)))
                self._check(self.lib.nds_get_hill_table_raw(self.h, _p(raw), raw.size))
                ck.update(hills_raw=raw, hills_n=n)
            ck["vr"] = self.get_vr()
            if self.cfg.mtd_adapsig_hist:
                ck["cv_history"], ck["cv_history_n"] = self.get_cv_history()
            n = C.c_int64()
            self._check(self.lib.nds_get_umb_vr(self.h, None, 0, C.byref(n)))
            if n.value:  # Harmonic.VR as of the last deposition (umbrella + metadynamics with frames)
                uv = np.empty(n.value)
                self._check(self.lib.nds_get_umb_vr(self.h, _p(uv), uv.size, C.byref(n)))
                ck["umb_vr"] = uv
        return ck

    def restore(self, ck):
        """Continue from :meth:`checkpoint`.  Runs ``begin()`` first (forces, constants), so an engine that records
        frames holds ONE pending frame labelled step 0 afterwards: callers that write dump files drop it with
        :meth:`discard_frames` (``NDRun.load_checkpoint`` does)."""
        self.set_state(ck["x"], ck["v"])
        self.begin()
        if "hills_raw" in ck:
            raw = _f64(ck["hills_raw"])
            self._check(self.lib.nds_set_hill_table_raw(self.h, int(ck["hills_n"]), _p(raw)))
            self.set_vr(ck["vr"])
        if "hills_heights" in ck:
            self.set_hills(ck["hills_centers"], ck["hills_heights"], sigma2=ck.get("hills_sigma2"))
            if "mtd_grid" in ck:
                self.set_mtd_grid(ck["mtd_grid"])
            self.set_vr(ck["vr"])
        if "mc_accepted" in ck:
            i64 = C.POINTER(C.c_int64)
            acc = np.ascontiguousarray(ck["mc_accepted"], dtype=np.int64)
            tm = _f64(ck["mc_time"]) if "mc_time" in ck else None
            dep = np.ascontiguousarray(ck["mc_dep"], dtype=np.int64) if "mc_dep" in ck else None
            hist = _f64(ck["mc_wl"]) if "mc_wl" in ck else None
            self._check(self.lib.nds_set_mc_state(self.h, acc.ctypes.data_as(i64), _p(tm),
                                                  None if dep is None else dep.ctypes.data_as(i64), _p(hist)))
            self._check(self.lib.nds_set_thermo(self.h, _p(_f64(ck["mc_thermo"]))))
        if "umb_vr" in ck:
            uv = _f64(ck["umb_vr"]).reshape(-1)
            self._check(self.lib.nds_set_umb_vr(self.h, _p(uv), uv.size))
        if "cv_history" in ck:  # after begin(), which appended a sample of its own
            self.set_cv_history(ck["cv_history"], int(ck["cv_history_n"]))
        if "alive" in ck:
            b = np.ascontiguousarray(ck["basin"], dtype=np.int32)
            e = np.ascontiguousarray(ck["exit_step"], dtype=np.int64)
            a = np.ascontiguousarray(ck["alive"], dtype=np.uint8)
            self._check(self.lib.nds_set_commit(self.h, b.ctypes.data_as(C.POINTER(C.c_int32)),
                                                e.ctypes.data_as(C.POINTER(C.c_int64)),
                                                a.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.set_clock(ck["step"], ck["time"], ck["dep_count"], ck["rescale_count"])

    def count_nonfinite(self):
        return int(self.lib.nds_count_nonfinite(self.h))

    # ---- committor
    def get_commit(self):
        basin = np.empty(self.M, dtype=np.int32)
        exit_step = np.empty(self.M, dtype=np.int64)
        self._check(self.lib.nds_get_commit(self.h, basin.ctypes.data_as(C.POINTER(C.c_int32)),
                                            exit_step.ctypes.data_as(C.POINTER(C.c_int64))))
        return basin, exit_step

    def get_mc(self):
        """accepted moves per walker (MC.accept_rate, engines/mc.py:85)."""
        acc = np.empty(self.M, dtype=np.int64)
        self._check(self.lib.nds_get_mc(self.h, acc.ctypes.data_as(C.POINTER(C.c_int64))))
        return acc

    def get_wl_histogram(self, walker=0):
        """WangLandau.hE[0] of one walker (engines/wang_landau.py:17)."""
        h = np.empty(_abi.NDS_WL_GRID)
        self._check(self.lib.nds_get_wl_histogram(self.h, int(walker), _p(h)))
        return h

    def set_wl_histogram(self, hE, walker=-1):
        """WangLandau.copy_hE with one basin (engines/wang_landau.py:24-34); after begin()."""
        h = _f64(hE, (_abi.NDS_WL_GRID,))
        self._check(self.lib.nds_set_wl_histogram(self.h, int(walker), _p(h)))

    def count_alive(self):
        return int(self.lib.nds_count_alive(self.h))

    # ---- frames
    @property
    def frames_f32(self):
        """True when the device frame ring holds float32 rows (fp32 engines)."""
        return bool(self.lib.nds_frames_f32(self.h)) if hasattr(self.lib, "nds_frames_f32") else False

    @property
    def frame_width(self):
        return int(self.lib.nds_frame_width(self.h))

    def frames_pending(self):
        return int(self.lib.nds_frames_pending(self.h))

    def drain_frames(self):
        """-> array ``[n_frames][frame_width][dump_walkers]`` (possibly empty)."""
        n = self.frames_pending()
        dw = int(min(self.cfg.dump_walkers, self.M))
        out = np.empty((n, self.frame_width, dw))
        got = C.c_int64()
        self._check(self.lib.nds_drain_frames(self.h, _p(out), n, C.byref(got)))
        return out[:got.value]

    def discard_frames(self):
        """Rewind the frame ring without reading it (nds_drain_frames with a NULL destination)."""
        self._check(self.lib.nds_drain_frames(self.h, None, 0, None))

    def get_histogram(self, clear=False):
        """Colvar histogram accumulated inside the step kernels since ``begin()``: ``counts[ny][nx]``."""
        ny = int(self.cfg.hist_n[1]) if self.cd > 1 else 1
        counts = np.zeros((ny, int(self.cfg.hist_n[0])), dtype=np.int64)
        self._check(self.lib.nds_get_histogram(self.h, counts.ctypes.data_as(C.POINTER(C.c_int64)), int(bool(clear))))
        return counts

    def frame_fields(self):
        nd, cd = self.ndim, self.cd
        names = ["step", "time"] + [f"x{d}" for d in range(nd)] + [f"v{d}" for d in range(nd)] + \
                [f"f{d}" for d in range(nd)] + [f"c{k}" for k in range(cd)] + \
                ["T", "pe", "ke", "biase", "totale", "VR"]
        # umbrella + metadynamics: Harmonic.VR of every restraint as of the last refresh (gaus.py:138-146)
        names += [f"VRu{u}" for u in range(self.frame_width - len(names))]
        return names

    # ---- plumbing / introspection
    def device_pointers(self):
        x, v, h = C.c_void_p(), C.c_void_p(), C.c_void_p()
        s = C.c_int64()
        self._check(self.lib.nds_device_pointers(self.h, C.byref(x), C.byref(v), C.byref(h), C.byref(s)))
        return dict(x=x.value, v=v.value, hills=h.value, hill_stride=s.value)

    @property
    def stream(self):
        return self.lib.nds_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.nds_launch_count(self.h))

    def last_run_ms(self):
        return float(self.lib.nds_last_run_ms(self.h))

    @property
    def kernel_name(self):
        return self.lib.nds_kernel_name(self.h).decode()


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId for :meth:`Engine.set_comm` (create on one rank, hand to all)."""
    lib = _abi.load_library()
    buf = C.create_string_buffer(128)
    if lib.nds_nccl_unique_id(buf):
        raise NdsError(lib.nds_last_error(None).decode())
    return buf.raw


def philox_kat(ctr, key, device=0):
    lib = _abi.load_library()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    if lib.nds_philox_kat(device, c, k, out):
        raise NdsError(lib.nds_last_error(None).decode())
    return list(out)


def noise_tail_counts(precision, seed, n_walkers, calls_per_walker, thresholds, device=0):
    """Census of the per-step Gaussian stream on the device: -> (counts of |z| > threshold, dict of moments)."""
    lib = _abi.load_library()
    thr = np.ascontiguousarray(thresholds, dtype=np.float64)
    counts = np.zeros(len(thr), dtype=np.uint64)
    stats = np.zeros(4)
    prec = _abi.NDS_F64 if precision == "f64" else _abi.NDS_F32
    if lib.nds_noise_tail_counts(device, prec, int(seed), int(n_walkers), int(calls_per_walker), _p(thr), len(thr),
                                 counts.ctypes.data_as(C.POINTER(C.c_uint64)), _p(stats)):
        raise NdsError(lib.nds_last_error(None).decode())
    n = 6.0 * n_walkers * calls_per_walker
    return counts.astype(np.int64), dict(n=n, mean=stats[0] / n, m2=stats[1] / n, m4=stats[2] / n, max=stats[3])


def map50to2d(x, jacobian=True, device=0):
    """``Map50to2d`` (``colvars/misc.py:19-156``) at configurations ``x[50][n]`` (or one ``x[50]``) on the device:
    -> ``dict(c=[2][n], J=[2][50][n])``.  The reference has no potential with ``ndim = 50``, so it cannot run a
    trajectory with this colvar; this evaluation is the whole surface it offers (``compute`` / ``jacobian``)."""
    lib = _abi.load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    single = x.ndim == 1
    if single:
        x = x.reshape(-1, 1)
    if x.ndim != 2 or x.shape[0] != 50:
        raise ValueError(f"Map50to2d takes x[50][n], got shape {x.shape}")
    n = x.shape[1]
    c = np.zeros((2, n))
    J = np.zeros((2, 50, n)) if jacobian else None
    if lib.nds_map50to2d(device, _p(x), n, _p(c), _p(J) if jacobian else None):
        raise NdsError(lib.nds_last_error(None).decode())
    if single:
        return dict(c=c[:, 0], J=None if J is None else J[:, :, 0])
    return dict(c=c, J=J)
