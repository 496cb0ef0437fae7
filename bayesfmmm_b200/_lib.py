"""Loader of libbfmmm_b200.so (ctypes).  Fails loudly when the CUDA extension is missing."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("BFMMM_LIB") or os.path.join(_HERE, "libbfmmm_b200.so")   # BFMMM_LIB: tuning builds only
_lib = None

dp = C.POINTER(C.c_double)


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("n", C.c_int32), ("K", C.c_int32), ("P", C.c_int32), ("M", C.c_int32),
                ("D", C.c_int32), ("device", C.c_int32), ("common_grid", C.c_int32), ("T", C.c_int64),
                ("off", C.POINTER(C.c_int64)), ("y", dp), ("B", dp), ("t", dp), ("degree", C.c_int32),
                ("n_internal", C.c_int32), ("internal_knots", dp), ("boundary", C.c_double * 2), ("X", dp),
                ("global_offset", C.c_int64)]


# every symbol include/bfmmm.h and include/bfmmm_debug.h declare
EXPORTS = [
    "bfmmm_create", "bfmmm_destroy", "bfmmm_last_error", "bfmmm_launch_count", "bfmmm_get_basis",
    "bfmmm_set_state", "bfmmm_get_state", "bfmmm_get_state_rows", "bfmmm_marginal_loglik", "bfmmm_cpo_reset", "bfmmm_cpo_accumulate", "bfmmm_cpo_get", "bfmmm_sigma_draw_async", "bfmmm_sigma_wait", "bfmmm_get_state_begin", "bfmmm_get_state_wait", "bfmmm_set_globals", "bfmmm_update_z", "bfmmm_update_chi",
    "bfmmm_ssr", "bfmmm_suffstats", "bfmmm_get_gram", "bfmmm_seed", "bfmmm_stats_buffer_dev",
    "bfmmm_update_z_async", "bfmmm_update_chi_async", "bfmmm_ssr_async", "bfmmm_suffstats_async",
    "bfmmm_read_stats", "bfmmm_clear_ssr_after", "bfmmm_sync", "bfmmm_stream", "bfmmm_engine_dims", "bfmmm_counts", "bfmmm_suffstats_ragged",
    # include/bfmmm_sampler.h
    "bfmmm_hyper_defaults", "bfmmm_sampler_create", "bfmmm_sampler_create_detached", "bfmmm_sampler_destroy", "bfmmm_sampler_device_resident", "bfmmm_sampler_set_allreduce", "bfmmm_nccl_unique_id", "bfmmm_sampler_enable_nccl", "bfmmm_nccl_destroy", "bfmmm_p2p_create", "bfmmm_sampler_enable_p2p", "bfmmm_p2p_destroy", "bfmmm_sampler_set_hband", "bfmmm_sampler_set_counts",
    "bfmmm_sampler_set", "bfmmm_sampler_get", "bfmmm_sampler_set_cov", "bfmmm_sampler_get_cov",
    "bfmmm_sampler_step", "bfmmm_sampler_run", "bfmmm_sampler_iteration", "bfmmm_sampler_last_accept",
    "bfmmm_sampler_tape", "bfmmm_sampler_tape_left", "bfmmm_sampler_set_tick", "bfmmm_sampler_tempered_transition",
    "bfmmm_sampler_run_mtt", "bfmmm_sampler_tt_trace", "bfmmm_sampler_record", "bfmmm_sampler_batches_written", "bfmmm_sampler_profile",
    # include/bfmmm_basis.h
    "bfmmm_bspline_basis", "bfmmm_tensor_bspline", "bfmmm_tensor_P", "bfmmm_get_P", "bfmmm_pmat_rw1",
    # include/bfmmm_io.h
    "bfmmm_save_mat_txt", "bfmmm_save_cube_txt", "bfmmm_save_field_cube_bin", "bfmmm_file_info", "bfmmm_load", "bfmmm_state_snapshot", "bfmmm_state_restore",
    "bfmmm_host_update_pi", "bfmmm_host_update_alpha3", "bfmmm_host_update_tau", "bfmmm_host_update_delta",
    "bfmmm_host_update_gamma", "bfmmm_host_update_A", "bfmmm_host_update_tau_eta", "bfmmm_host_update_delta_xi",
    "bfmmm_host_update_gamma_xi", "bfmmm_host_update_A_xi", "bfmmm_host_update_phi", "bfmmm_host_update_nu",
    "bfmmm_host_update_eta", "bfmmm_host_update_xi", "bfmmm_host_update_sigma",
    # include/bfmmm_post.h
    "bfmmm_quantiles",
    "bfmmm_debug_enable_acc", "bfmmm_debug_get_acc", "bfmmm_debug_update_z_rng", "bfmmm_debug_update_chi_rng",
    "bfmmm_debug_get_cache", "bfmmm_debug_fastmath", "bfmmm_debug_moments_valid", "bfmmm_debug_z_propose",
]


def library_path() -> str:
    return _LIB


def source_hash() -> str:
    """md5 over the library's sources (csrc/ and include/): identifies the build independently of when it was compiled."""
    import hashlib
    h = hashlib.md5()
    for d in (os.path.join(_HERE, "csrc"), os.path.join(os.path.dirname(_HERE), "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


# translation units without device code of the sweep: the driver loop, the all-reduce hooks, file I/O, host linear algebra
_HOST_ONLY = {"host_sampler.cu", "p2p_hook.cu", "nccl_hook.cu", "host_linalg.cpp", "arma_io.cu", "basis_host.cu"}


def kernel_source_hash() -> str:
    """md5 over the device-code sources (csrc/ minus the host-only translation units): the key of the committed ncu
    traffic evidence -- a kernel's DRAM bytes do not depend on the host loop around it."""
    import hashlib
    h = hashlib.md5()
    d = os.path.join(_HERE, "csrc")
    for f in sorted(os.listdir(d)):
        if f in _HOST_ONLY:
            continue
        if f.endswith((".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def build_library(jobs: int = 8, extra: str = "") -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}", "-s"]
    if extra:
        cmd.append(f"EXTRA={extra}")
    subprocess.check_call(cmd)
    return _LIB


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise RuntimeError(
            f"{_LIB} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the sampler hot path)")
    lib = C.CDLL(_LIB)
    lib.bfmmm_last_error.restype = C.c_char_p
    lib.bfmmm_launch_count.restype = C.c_int64
    lib.bfmmm_stream.restype = C.c_void_p
    lib.bfmmm_stream.argtypes = [C.c_void_p]
    lib.bfmmm_destroy.restype = None
    lib.bfmmm_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib
