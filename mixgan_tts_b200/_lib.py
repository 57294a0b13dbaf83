"""ctypes binding of ``libmixgan_b200.so`` (the C ABI declared in ``include/mixgan_b200.h``).

The product path has no CPU fallback: if the shared library is missing, or a call
returns a non-zero code, a ``RuntimeError`` / ``ValueError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmixgan_b200.so")

PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 3
PACK_FP32_TABLES = 2       # mgb_pack_weights only: the fp32 buffer with just the per-utterance table weights
E_ARG, E_ARCH, E_WORKSPACE, E_CUDA, E_UNSUPPORTED = -1, -2, -3, -4, -5


class ModelDims(C.Structure):
    _fields_ = [("n_mel", C.c_int32), ("channels", C.c_int32), ("d_encoder", C.c_int32),
                ("layers", C.c_int32), ("multi_speaker", C.c_int32)]


_P, _I, _Z, _F = C.c_void_p, C.c_int, C.c_size_t, C.c_float
_D = C.POINTER(ModelDims)

# name -> (restype, argtypes): every symbol include/mixgan_b200.h declares
SIGNATURES = {
    "mgb_abi_version": (_I, []),
    "mgb_last_error": (C.c_char_p, []),
    "mgb_launch_count": (C.c_longlong, []),
    "mgb_note_launches": (None, [C.c_longlong]),
    "mgb_profile_enable": (None, [_I]),
    "mgb_profile_collect": (_I, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "mgb_profile_read_stamps": (_I, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float)]),
    "mgb_debug_status": (_I, [_D, _I, _I, _I, _P, C.POINTER(C.c_int)]),
    "mgb_device_check": (_I, [_I]),
    "mgb_flat_weight_count": (_Z, [_D]),
    "mgb_packed_bytes": (_Z, [_D, _I]),
    "mgb_pack_weights": (_I, [_D, _I, _P, _P, _Z, _P]),
    "mgb_workspace_bytes": (_Z, [_D, _I, _I, _I, _I]),
    "mgb_denoiser_forward": (_I, [_D, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "mgb_reverse_step": (_I, [_D, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _P, _Z, _P]),
    "mgb_sample": (_I, [_D, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "mgb_train_saved_bytes": (_Z, [_D, _I, _I, _I]),
    "mgb_train_workspace_bytes": (_Z, [_D, _I, _I, _I]),
    "mgb_train_debug_status": (_I, [_D, _I, _I, _P, C.POINTER(C.c_int)]),
    "mgb_train_segments": (_I, [_D]),
    "mgb_train_segment_range": (_I, [_D, _I, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mgb_denoiser_train_forward": (_I, [_D, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _I, _I, _P, _Z, _P]),
    "mgb_denoiser_backward": (_I, [_D, _I, _P, _P, _Z, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "mgb_pack_cond": (_I, [_D, _I, _P, _I, _I, _P, _Z, _P]),
    "mgb_shallow_start": (_I, [_P, _P, _P, _P, _F, _F, _P, _P, _I, _I, _I, _P]),
    "mgb_denorm_mask": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "mgb_train_diffuse": (_I, [_P] * 11 + [_I] * 4 + [_P]),
    "mgb_train_posterior": (_I, [_P] * 6 + [_I] + [_P, _P] + [_I] * 4 + [_P]),
    "mgb_train_posterior_backward": (_I, [_P] * 6 + [_I] + [_P] + [_I] * 4 + [_P]),
    "mgb_conv1d_out_len": (_I, [_I, _I, _I]),
    "mgb_conv1d_workspace_bytes": (_Z, [_I] * 6),
    "mgb_conv1d_forward": (_I, [_P] * 6 + [_I] * 7 + [_P, _Z, _P]),
    "mgb_conv1d_backward": (_I, [_P] * 10 + [_I] * 7 + [_P, _Z, _P]),
    "mgb_step_embedding": (_I, [_P, _P, _I, _I, _P]),
    "mgb_durations_from_log": (_I, [_P, _F, _P, _I, _P]),
    "mgb_length_regulate": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "mgb_length_regulate_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "mgb_mask_from_lengths": (_I, [_P, _P, _I, _I, _P]),
    "mgb_auxdec_flat_count": (_Z, [_P]),
    "mgb_auxdec_packed_bytes": (_Z, [_P]),
    "mgb_auxdec_workspace_bytes": (_Z, [_P, _I, _I]),
    "mgb_auxdec_pack": (_I, [_P, _P, _P, _Z, _P]),
    "mgb_auxdec_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "mgb_auxdec_debug_status": (_I, [_P, _I, _I, _P, _P]),
    "mgb_hifigan_flat_count": (_Z, [_P]),
    "mgb_hifigan_packed_bytes": (_Z, [_P]),
    "mgb_hifigan_hop": (_I, [_P]),
    "mgb_hifigan_workspace_bytes": (_Z, [_P, _I, _I]),
    "mgb_hifigan_pack": (_I, [_P, _P, _P, _Z, _P]),
    "mgb_hifigan_forward": (_I, [_P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "mgb_hifigan_debug_status": (_I, [_P, _I, _I, _P, _P]),
}

# include/mixgan_b200_probe.h: exported by the test-only debug library (libmixgan_b200_dbg.so), never by the product one
PROBE_SIGNATURES = {
    "mgb_probe_umma": (_I, [_P, _I, _P, _I] + [_I] * 11 + [_P, _P, _P]),
    "mgb_probe_umma_2cta": (_I, [_P, _I, _P, _I] + [_I] * 8 + [_P, _P, _P]),
    "mgb_probe_bulk_rate": (_I, [_P, C.c_longlong, _I, _I, _I, _I, _I, _P, _P, _P]),
    "mgb_probe_umma_rate": (_I, [_I] * 13 + [_P, _P, _P]),
    "mgb_probe_umma_rate_data": (_I, [_I] * 15 + [_P, _P, _P]),
}

_lib = None
_dbg = None
LIB_DBG_PATH = os.path.join(HERE, "libmixgan_b200_dbg.so")


def load_debug():
    """The test/diagnostic superset library (product ABI + probes + in-kernel profiling).  Tests and scripts only."""
    global _dbg
    if _dbg is not None:
        return _dbg
    if not os.path.exists(LIB_DBG_PATH):
        raise RuntimeError(f"{LIB_DBG_PATH} is missing: build it with `python -m mixgan_tts_b200.build`")
    lib = C.CDLL(LIB_DBG_PATH)
    for name, (res, args) in {**SIGNATURES, **PROBE_SIGNATURES}.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _dbg = lib
    return lib


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    # diagnostics scripts (scripts/gpu_prof.sh) run the product ABI out of the debug build: MIXGAN_B200_USE_DEBUG_LIB=1
    path = LIB_DBG_PATH if os.environ.get("MIXGAN_B200_USE_DEBUG_LIB") == "1" else LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m mixgan_tts_b200.build` "
            "(there is no CPU fallback for this path)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.mgb_abi_version() != 2:
        raise RuntimeError("libmixgan_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str, lib=None):
    if rc == 0:
        return
    msg = (lib or load()).mgb_last_error().decode(errors="replace")
    err = ValueError if rc in (E_ARG, E_WORKSPACE) else RuntimeError
    raise err(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
