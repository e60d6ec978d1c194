"""ctypes mirror of ``include/ndsb200.h`` and the loader of the CUDA shared library.

There is no CPU fallback: :func:`load_library` raises if ``libndsb200.so`` has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C ndsimulator_b200/csrc``).
"""
import ctypes as C
import os

NDS_ABI_VERSION = 3
NDS_MAX_DIM = 5
NDS_MAX_CV = 5
NDS_MAX_UMB = 4
NDS_MAX_BASINS = 8

NDS_F32, NDS_F64 = 0, 1
NDS_METHOD_MD, NDS_METHOD_COMMITTOR, NDS_METHOD_MC, NDS_METHOD_WL = 0, 1, 2, 3
NDS_WL_GRID = 101
NDS_INT_LANGEVIN, NDS_INT_LANGEVIN2, NDS_INT_NVE, NDS_INT_RESCALE, NDS_INT_MINIMIZE = 0, 1, 2, 3, 4
NDS_NOISE_PHILOX, NDS_NOISE_EXTERNAL = 0, 1

INTEGRATORS = {"langevin": NDS_INT_LANGEVIN, "2nd-langevin": NDS_INT_LANGEVIN2,
               "nve": NDS_INT_NVE, "rescale": NDS_INT_RESCALE, "minimize": NDS_INT_MINIMIZE}

# class names of ndsimulator/potentials (reference potentials/__init__.py) -> (id, ndim, require_colvar)
POTENTIALS = {
    "Mueller2d": (0, 2, False),
    "Mueller5d": (1, 5, True),
    "Mueller5dshl": (2, 5, True),
    "DoubleWell1d": (3, 1, False),
    "DoubleWell": (4, 2, False),
    "ThreeHole2d": (5, 2, True),
    "ThreeHole5d": (6, 5, True),
    "Gaussian2d": (7, 2, False),
    "Gaussian": (8, 2, False),
    "Flat1d": (9, 1, False),
    "Flat2d": (10, 2, False),
    "HarmonicQuartic": (11, None, False),  # extension (no reference class): ndim 1, 2 or 5
}

# class names of ndsimulator/colvars -> (id, required ndim or None, colvardim or None(=ndim))
COLVARS = {
    "Original": (0, None, None),
    "ExtractX": (1, 2, 1),
    "ExtractY": (2, 2, 1),
    "XmY": (3, 2, 1),
    "XpY": (4, 2, 1),
    "Rotate2d": (5, 2, 2),
    "Proj2dto1d": (6, 2, 1),
    "Proj5dto1dx0": (7, 5, 1),
    "Proj5dto1dx2": (8, 5, 1),
    "Proj5dto1d": (9, 5, 1),
    "Proj5dto2d": (10, 5, 2),
    "Proj5dto2dv2": (11, 5, 2),
    "Path": (12, None, 1),
    "Path5dto2d": (13, 5, 1),
}


class NdsConfig(C.Structure):
    """``struct nds_config`` (include/ndsb200.h)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32),
        ("ndim", C.c_int32), ("method", C.c_int32), ("integrator", C.c_int32),
        ("potential", C.c_int32), ("colvar", C.c_int32), ("true_colvar", C.c_int32),
        ("noise_mode", C.c_int32), ("steps_per_launch", C.c_int32), ("reserved0", C.c_int32),
        ("n_walkers", C.c_int64), ("walker_offset", C.c_int64), ("n_walkers_global", C.c_int64),
        ("seed", C.c_uint64),
        ("temperature", C.c_double), ("initial_temperature", C.c_double), ("mass", C.c_double),
        ("dt", C.c_double), ("gamma", C.c_double), ("rescale_freq", C.c_double),
        ("pot_params", C.c_double * 16),
        ("n_umb", C.c_int32), ("umb_per_walker", C.c_int32),
        ("umb_k", (C.c_double * NDS_MAX_CV) * NDS_MAX_UMB),
        ("umb_r0", (C.c_double * NDS_MAX_CV) * NDS_MAX_UMB),
        ("mtd_enabled", C.c_int32), ("mtd_shared", C.c_int32),
        ("mtd_w", C.c_double), ("mtd_sigma", C.c_double * NDS_MAX_CV),
        ("mtd_dep_freq", C.c_double), ("mtd_biasf", C.c_double), ("mtd_capacity", C.c_int64),
        ("mtd_adapsig_geo", C.c_int32), ("reserved3", C.c_int32),
        ("pe_enabled", C.c_int32), ("reserved1", C.c_int32), ("pe_thred", C.c_double),
        ("n_basins", C.c_int32), ("reserved2", C.c_int32), ("diffusion_time", C.c_int64),
        ("criteria", ((C.c_double * 2) * NDS_MAX_CV) * NDS_MAX_BASINS),
        ("dump_freq", C.c_int64), ("dump_walkers", C.c_int64), ("dump_capacity", C.c_int64),
        ("stat_freq", C.c_int64),
        ("emin", C.c_double), ("min_gtol", C.c_double), ("min_maxiter", C.c_int64),
        ("mc_bounds", C.c_double * NDS_MAX_DIM), ("mc_global_bounds", C.c_double * NDS_MAX_DIM),
        ("mc_propose_all", C.c_int32), ("mc_periodic", C.c_int32),
        ("mtd_grid", C.c_int32), ("mtd_grid_n", C.c_int32 * 2), ("reserved4", C.c_int32),
        ("mtd_grid_lo", C.c_double * 2), ("mtd_grid_hi", C.c_double * 2),
        ("mtd_adapsig_hist", C.c_int32), ("reserved5", C.c_int32),
        ("hist_n", C.c_int32 * 2), ("hist_lo", C.c_double * 2), ("hist_hi", C.c_double * 2),
    ]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64)

_HERE = os.path.dirname(os.path.abspath(__file__))
# NDSB200_LIB: another build of the same library (kernel experiments under profiles/experiments/); default: in-tree
LIB_PATH = os.environ.get("NDSB200_LIB") or os.path.join(_HERE, "csrc", "libndsb200.so")
_lib = None

_dp = C.POINTER(C.c_double)
_PROTOTYPES = {
    "nds_last_error": (C.c_char_p, [C.c_void_p]),
    "nds_config_defaults": (None, [C.POINTER(NdsConfig)]),
    "nds_colvar_dim": (C.c_int, [C.c_int, C.c_int]),
    "nds_create": (C.c_int, [C.POINTER(NdsConfig), C.POINTER(C.c_void_p)]),
    "nds_destroy": (None, [C.c_void_p]),
    "nds_set_state": (C.c_int, [C.c_void_p, _dp, _dp]),
    "nds_set_walker_bias": (C.c_int, [C.c_void_p, _dp, _dp]),
    "nds_set_path": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _dp, _dp, C.c_double]),
    "nds_set_noise": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_int64]),
    "nds_begin": (C.c_int, [C.c_void_p]),
    "nds_run": (C.c_int, [C.c_void_p, C.c_int64]),
    "nds_sync": (C.c_int, [C.c_void_p]),
    "nds_get_state": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "nds_set_state_f32": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "nds_get_state_f32": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_float)] * 6),
    "nds_step": (C.c_int64, [C.c_void_p]),
    "nds_time": (C.c_double, [C.c_void_p]),
    "nds_eval_at": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp]),
    "nds_get_hills": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), _dp, _dp, C.c_int64]),
    "nds_set_hills": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _dp, _dp]),
    "nds_hill_table_size": (C.c_int64, [C.c_void_p, C.c_int64]),
    "nds_get_hill_table_raw": (C.c_int, [C.c_void_p, _dp, C.c_int64]),
    "nds_set_hill_table_raw": (C.c_int, [C.c_void_p, C.c_int64, _dp]),
    "nds_get_hill_sigma2": (C.c_int, [C.c_void_p, C.c_int64, _dp, C.c_int64]),
    "nds_set_hill_sigma2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _dp]),
    "nds_get_vr": (C.c_int, [C.c_void_p, _dp]),
    "nds_set_exchange": (C.c_int, [C.c_void_p, EXCHANGE_FN, C.c_void_p]),
    "nds_hills_ipc_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nds_open_peer_tables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nds_set_peer_tables": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "nds_set_peer_barrier": (C.c_int, [C.c_void_p, C.c_int]),
    "nds_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "nds_set_comm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nds_exchange_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "nds_table_checksum": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]),
    "nds_get_mtd_grid": (C.c_int, [C.c_void_p, _dp, C.c_int64]),
    "nds_set_mtd_grid": (C.c_int, [C.c_void_p, _dp, C.c_int64]),
    "nds_mtd_grid_outside": (C.c_int64, [C.c_void_p]),
    "nds_get_mc_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "nds_set_mc_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_double)]),
    "nds_set_thermo": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "nds_get_umb_vr": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64)]),
    "nds_set_umb_vr": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int64]),
    "nds_get_cv_history": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nds_set_cv_history": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int64]),
    "nds_get_histogram": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "nds_frames_f32": (C.c_int, [C.c_void_p]),
    "nds_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "nds_wait": (C.c_int, [C.c_void_p]),
    "nds_get_commit": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "nds_get_alive": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)]),
    "nds_set_commit": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]),
    "nds_count_alive": (C.c_int64, [C.c_void_p]),
    "nds_get_mc": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "nds_get_wl_histogram": (C.c_int, [C.c_void_p, C.c_int64, _dp]),
    "nds_set_wl_histogram": (C.c_int, [C.c_void_p, C.c_int64, _dp]),
    "nds_get_clock": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nds_set_clock": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_int64, C.c_int64]),
    "nds_set_vr": (C.c_int, [C.c_void_p, _dp]),
    "nds_count_nonfinite": (C.c_int64, [C.c_void_p]),
    "nds_frame_width": (C.c_int, [C.c_void_p]),
    "nds_frames_pending": (C.c_int64, [C.c_void_p]),
    "nds_drain_frames": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.POINTER(C.c_int64)]),
    "nds_fes_grid": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int64,
                               C.c_double, C.c_double, C.c_int64, _dp]),
    "nds_colvar_histogram": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_double,
                                       C.c_double, C.c_int64, C.POINTER(C.c_int64)]),
    "nds_device_pointers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "nds_stream": (C.c_void_p, [C.c_void_p]),
    "nds_n_hills": (C.c_int64, [C.c_void_p, C.c_int64]),
    "nds_launch_count": (C.c_int64, [C.c_void_p]),
    "nds_last_run_ms": (C.c_double, [C.c_void_p]),
    "nds_kernel_name": (C.c_char_p, [C.c_void_p]),
    "nds_philox_kat": (C.c_int, [C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint32)]),
    "nds_noise_tail_counts": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.POINTER(C.c_double),
                                        C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "nds_map50to2d": (C.c_int, [C.c_int, _dp, C.c_int64, _dp, _dp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def load_library(path: str = None):
    """dlopen ``libndsb200.so`` and attach prototypes.  Raises ``RuntimeError`` when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension has not been built (run __graft_entry__.build()). "
            "ndsimulator_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
