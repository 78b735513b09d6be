"""ctypes binding of libbqo_sm100.so (declared in include/bqo_sm100.h).

There is no CPU fallback: importing this module without the built library raises, and every
compute call requires CUDA tensors.  PyTorch only owns the device buffers and the stream.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from ._build import LIB_PATH

c_double_p = C.c_void_p  # device pointers are passed as raw addresses
c_int_p = C.c_void_p

KIND_MATERN52 = 0
KIND_SCALED_MATERN52 = 1
KIND_PRODUCT_TASKS = 2
MAX_DIM = 16


class KernelDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dim", C.c_int32),
        ("dim_m", C.c_int32),
        ("n_tasks", C.c_int32),
        ("same_corr", C.c_int32),
        ("n_params", C.c_int32),
    ]


class StageC(C.Structure):
    _fields_ = [
        ("desc", KernelDesc),
        ("n", C.c_int32),
        ("S", C.c_int32),
        ("dx", C.c_int32),
        ("dw", C.c_int32),
        ("n_units", C.c_int32),
        ("C", C.c_int32),
        ("M", C.c_int32),
        ("n_wsets", C.c_int32),
        ("hyp", C.c_void_p),
        ("Xs", C.c_void_p),
        ("tid", C.c_void_p),
        ("alpha", C.c_void_p),
        ("qt", C.c_void_p),
        ("R", C.c_void_p),
        ("den", C.c_void_p),
        ("beta4", C.c_void_p),
        ("cand", C.c_void_p),
        ("unit_k", C.c_void_p),
        ("unit_c", C.c_void_p),
        ("wsets", C.c_void_p),
        ("lower", C.c_void_p),
        ("upper", C.c_void_p),
        ("wdist", C.c_void_p),
    ]


class InnerOpt(C.Structure):
    _fields_ = [
        ("n_z", C.c_int32),
        ("n_starts", C.c_int32),
        ("maxiter", C.c_int32),
        ("max_ls", C.c_int32),
        ("factr", C.c_double),
        ("pgtol", C.c_double),
        ("use_vertices", C.c_int32),
        ("z_per_cta", C.c_int32),
        ("starts_per_hyper", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_D = C.c_double
_KD = C.POINTER(KernelDesc)
_SC = C.POINTER(StageC)
_IO = C.POINTER(InnerOpt)

# name -> (restype, argtypes); mirrors include/bqo_sm100.h one to one
SIGNATURES = {
    "bqo_abi_version": (C.c_int, []),
    "bqo_set_device": (C.c_int, [C.c_int]),
    "bqo_launch_count": (_L, []),
    "bqo_hyper_stride": (_L, [_KD]),
    "bqo_fp64_peak_probe_workspace_bytes": (C.c_int64, []),
    "bqo_fp64_peak_probe": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(_D), C.POINTER(_D)]),
    "bqo_hbm_copy_probe": (C.c_int, [_P, _P, _L, C.c_int, C.POINTER(_D)]),
    "bqo_hyper_prepare": (C.c_int, [_KD, _P, _I, _P, _P]),
    "bqo_scale_points": (C.c_int, [_KD, _P, _I, _P, _I, _I, _P, _P, _P]),
    "bqo_kernel_assemble_batched": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _I, _P, _P, _P]),
    "bqo_kernel_grad_params_batched": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _P]),
    "bqo_cross_cov_batched": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _I, _I, _P, _P, _P]),
    "bqo_cross_cov_hessian_batched": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _I, _P, _P]),
    "bqo_potrf_batched": (C.c_int, [_P, _I, _I, _P, _P, _P]),
    "bqo_potrf_workspace_bytes": (_L, [_I, _I]),
    "bqo_potrf_set_multi_launch": (C.c_int, [C.c_int]),
    "bqo_potrf_profile": (C.c_int, [_P, _I, _I, _P, _P, _P, _P]),
    "bqo_potrs_batched": (C.c_int, [_P, _I, _I, _P, _I, _P]),
    "bqo_chol_cov_batched": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "bqo_chol_cov_retry": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _D, _P, _P, _P, _P, _P]),
    "bqo_chol_append_batched": (C.c_int, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "bqo_alpha_batched": (C.c_int, [_P, _L, _P, _I, _I, _P, _P, _P]),
    "bqo_alpha_from_inverse_batched": (C.c_int, [_P, _L, _P, _I, _I, _P, _P, _P, _P]),
    "bqo_loglik_batched": (C.c_int, [_P, _L, _P, _P, _P, _I, _I, _P, _P]),
    "bqo_loglik_grad_workspace_bytes": (_L, [_I, _I]),
    "bqo_loglik_grad_batched": (C.c_int, [_KD, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "bqo_posterior_batched": (
        C.c_int,
        [_KD, _P, _I, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    ),
    "bqo_ei_batched": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "bqo_task_quadrature_weights": (C.c_int, [_KD, _P, _I, _P, _P, _P]),
    "bqo_candidate_setup_batched": (
        C.c_int,
        [_KD, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    ),
    "bqo_weight_alpha": (C.c_int, [_KD, _P, _P, _P, _I, _I, _P, _P]),
    "bqo_wdist_precompute": (C.c_int, [_KD, _P, _I, _P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "bqo_sample_ab_batched": (C.c_int, [_SC, _P, _P, _P, _I, _P, _P]),
    "bqo_sample_ab_hess_batched": (C.c_int, [_SC, _P, _P, _P, _I, _P, _P]),
    "bqo_inner_maximize_batched": (C.c_int, [_SC, _IO, _P, _P, _P, _P, _P, _P]),
    "bqo_grad_vector_b_batched": (C.c_int, [_SC, _P, _I, _P, _P, _P, _P]),
    "bqo_acq_reduce": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "bqo_kg_workspace_bytes": (_L, [_I, _I]),
    "bqo_kg_discretized_batched": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "bqo_adam_step_batched": (
        C.c_int,
        [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _D, _D, _D, _D, _P],
    ),
}

_lib = None


class BqoLibraryError(RuntimeError):
    pass


def load(path: Path | None = None) -> C.CDLL:
    """Load the shared library and attach the declared signatures.  Fails loudly if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    import os
    if path is None and os.environ.get("BQO_LIB_PATH"):
        p = Path(os.environ["BQO_LIB_PATH"])  # tuning variants built by _build.build_variant
    else:
        p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise BqoLibraryError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % p
        )
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.bqo_abi_version() != 1:
        raise BqoLibraryError("ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise BqoLibraryError("%s: bad argument #%d" % (what, -rc))
    raise BqoLibraryError("%s: CUDA error %d" % (what, rc))


def ptr(t) -> int | None:
    """Device address of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise BqoLibraryError("expected a CUDA tensor (no CPU path exists)")
    if not t.is_contiguous():
        raise BqoLibraryError("expected a contiguous tensor")
    return t.data_ptr()
