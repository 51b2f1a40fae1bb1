"""ctypes binding of the C ABI declared in include/glia_rd.h.

The product library is ``glia_b200/lib/libglia_rd.so`` (nvcc, sm_100a).  There is no
CPU fallback: if the library or a CUDA device is missing, loading / handle creation
raises.  (tests/emu builds the *same sources* against a thread-per-CUDA-thread SIMT
emulator to debug kernel index logic on GPU-less machines; that build is only ever
loaded by tests through ``load_library(path)`` and is not part of the package.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libglia_rd.so")

# name -> (restype, argtypes); every symbol include/glia_rd.h declares
_P = C.c_void_p
_I = C.c_int
_D = C.c_double
SIGNATURES = {
    "glia_rd_abi_version": (_I, []),
    "glia_rd_build_info": (C.c_char_p, []),
    "glia_rd_create": (_I, [C.POINTER(_P), C.POINTER(_I), _I, _I, _D]),
    "glia_rd_create_slab": (_I, [C.POINTER(_P), C.POINTER(_I), _I, _I, _D, _I, _I]),
    "glia_rd_create_batch": (_I, [C.POINTER(_P), C.POINTER(_I), _I, _I, _D, _I]),
    "glia_rd_batch_size": (_I, [_P, C.POINTER(_I)]),
    "glia_rd_batch_iterations": (_I, [_P, C.POINTER(_I), _I]),
    "glia_rd_set_coefficients_batch": (_I, [_P, _P, _P, _P, C.POINTER(_D), _D, _D, _D, C.POINTER(_D), _D, _D]),
    "glia_rd_ipc_export": (_I, [_P, _I, _P]),
    "glia_rd_ipc_connect": (_I, [_P, _I, _P]),
    "glia_rd_ipc_disconnect": (_I, [_P, _I]),
    "glia_rd_wait_stream": (_I, [_P, _P]),
    "glia_rd_set_splitting_order": (_I, [_P, _I]),
    "glia_rd_set_two_snapshot": (_I, [_P, _P, _P]),
    "glia_rd_destroy": (_I, [_P]),
    "glia_rd_last_error": (C.c_char_p, [_P]),
    "glia_rd_stream": (_P, [_P]),
    "glia_rd_launch_count": (C.c_longlong, [_P]),
    "glia_rd_fft_r2c": (_I, [_P, _P, _P]),
    "glia_rd_fft_c2r": (_I, [_P, _P, _P]),
    "glia_rd_gradient": (_I, [_P, _P, _P, _P, _P, _I]),
    "glia_rd_divergence": (_I, [_P, _P, _P, _P, _P]),
    "glia_rd_set_diffusion": (_I, [_P, _P, C.POINTER(_D), _D]),
    "glia_rd_set_diffusion_tissue": (_I, [_P, _P, _P, _P, _D, _D, _D, _D]),
    "glia_rd_set_secondary_k": (_I, [_P, _P]),
    "glia_rd_set_reaction": (_I, [_P, _P]),
    "glia_rd_set_reaction_tissue": (_I, [_P, _P, _P, _P, _D, _D, _D]),
    "glia_rd_update_reac_diff": (_I, [_P, _P, _P, _P, _P, _D, _D, _D, _D]),
    "glia_rd_apply_D": (_I, [_P, _P, _P, _I]),
    "glia_rd_prec_factor": (_I, [_P]),
    "glia_rd_diffusion_solve": (_I, [_P, _P, _D, C.POINTER(_I)]),
    "glia_rd_set_ksp_tolerances": (_I, [_P, _D, _D, _D, _I]),
    "glia_rd_resize_history": (_I, [_P, _I, _D]),
    "glia_rd_history": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "glia_rd_reaction": (_I, [_P, _P, _P, _D]),
    "glia_rd_solve_state": (_I, [_P, _P, _P, _I, C.POINTER(_I)]),
    "glia_rd_solve_adjoint": (_I, [_P, _P, _P, _I, _I, C.POINTER(_I)]),
    "glia_rd_grad_kappa_rho": (_I, [_P, _P, _P, _P, C.POINTER(_D)]),
    "glia_rd_set_secondary_tissue": (_I, [_P, _P, _P, _P, _D, _D, _D]),
    "glia_rd_objective_gradient": (_I, [_P, _P, _P, _P, _D, _P, _P, _P, C.POINTER(_D), _P, C.POINTER(_D),
                                        C.POINTER(_I)]),
    "glia_rd_hessian_matvec": (_I, [_P, _P, _P, _D, _I, _P, _P, _P, _P, C.POINTER(_D), C.POINTER(_I)]),
    "glia_rd_smooth": (_I, [_P, _P, _P, _D]),
    "glia_rd_mat_prop": (_I, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(_D)]),
    "glia_rd_phi_set": (_I, [_P, _I, C.POINTER(_D), _D, _P, _D]),
    "glia_rd_phi_apply": (_I, [_P, _P, C.POINTER(_D)]),
    "glia_rd_phi_apply_transpose": (_I, [_P, C.POINTER(_D), _P]),
    "glia_rd_data_in": (_I, [_P, C.c_char_p, _P]),
    "glia_rd_data_out": (_I, [_P, C.c_char_p, _P]),
    "glia_rd_split_segmentation": (_I, [_P, _P, C.POINTER(_I), _P, _P, _P, _P]),
    "glia_rd_probe_xsweep": (_I, [_P, _I, _I, _I, C.POINTER(_D)]),
    "glia_rd_profile_begin": (_I, [_P]),
    "glia_rd_profile_end": (_I, [_P, C.c_char_p, _I]),
    "glia_rd_timer_start": (_I, [_P]),
    "glia_rd_timer_stop_ms": (_I, [_P, C.POINTER(_D)]),
    "glia_rd_forward_adjoint": (_I, [_P, _P, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)]),
    "glia_rd_forward_adjoint_host": (_I, [_P, _P, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)]),
}

_libs = {}


class GliaRdError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen the C-ABI library and bind every declared symbol (raises if any is missing)."""
    path = path or os.environ.get("GLIA_RD_LIB", DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise GliaRdError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). glia_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib


def declared_symbols(header_path: str) -> list[str]:
    """Names of the functions declared in a C header (used by the ABI test)."""
    import re
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(glia_rd_[A-Za-z0-9_]+)\s*\(", txt)))
