"""ctypes binding of libcsmae_b200.so (C-ABI declared in include/csmae_b200.h).

This is the only place the Python host talks to native code: torch supplies device memory
(`tensor.data_ptr()`) and the current CUDA stream, every op below is one `extern "C"` call.  There
is no fallback: if the library is missing or the device is not sm_100, importing/using the ops
raises.
"""
import ctypes
import os
import threading

import torch

from . import build as _build

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_ll = ctypes.c_longlong

EPI_BF16, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_F32 = 0, 1, 2, 3, 5

# name -> argtypes (every function returns int except csm_last_error); mirrors include/csmae_b200.h
_P, _I, _F, _L = _c_void_p, _c_int, _c_float, _c_ll
SIGNATURES = {
    "csm_version": [],
    "csm_device_check": [_I],
    "csm_linear_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_linear_dgrad": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_linear_wgrad": [_P, _P, _P, _I, _I, _I, _I, _P],
    "csm_gemm_workspace_bytes": [_I],
    "csm_gemm_set_workspace": [_P, _L],
    "csm_colsum_bf16": [_P, _P, _I, _I, _I, _I, _P],
    "csm_random_masking": [_P, _I, _I, _I, _P, _P, _P, _P],
    "csm_resized_crop": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "csm_patch_gather": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "csm_encoder_assemble": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_decoder_assemble": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_decoder_assemble_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_encoder_out_grad": [_P, _P, _P, _P, _I, _I, _I, _P],
    "csm_cls_grad": [_P, _P, _I, _I, _I, _P],
    "csm_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "csm_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "csm_cast_multi": [_P, _I, _I, _P],
    "csm_cast_f32_bf16": [_P, _P, _L, _P],
    "csm_attention_fwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "csm_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_recon_loss_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "csm_recon_loss_bwd": [_P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _I, _P],
    "csm_cross_mse_fwd": [_P, _P, _P, _I, _I, _I, _P],
    "csm_cross_mse_bwd": [_P, _P, _P, _P, _P, _F, _I, _I, _I, _P],
    "csm_bn_patch_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P],
    "csm_bn_patch_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csm_ntxent_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P],
    "csm_ntxent_bwd": [_P, _P, _P, _P, _P, _I, _I, _F, _F, _P],
    "csm_adamw_multi": [_P, _P, _I, _P, _I, _P],
    "csm_grad_stats_f32": [_P, _L, _P, _P, _I, _P],
    "csm_sincos_pos_embed": [_P, _I, _I, _I, _P],
    "csm_loss_finalize": [_P, _P, _P, _I, _P],
    "csm_zero_async": [_P, _L, _P],
    "csm_amp_update": [_P, _P, _P, _P, _F, _F, _F, _I, _P],
    "csm_sumsq_f32": [_P, _L, _P, _I, _P],
}

# development-only exports (not part of the C-ABI in include/csmae_b200.h): the mma.sync attention kernels kept as the
# A/B baseline and the forward variant switch, used by tools/attn_tc_check.py
DEV_SIGNATURES = {
    "csm_attention_fwd_tc": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "csm_attention_fwd_legacy": [_P, _P, _P, _I, _I, _I, _I, _P],
    "csm_attention_bwd_legacy": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
}

_lib = None
_lock = threading.Lock()
_sm_count = {}
launch_count = 0   # kernels launched through the C-ABI so far (bench.py reports the delta)
KERNELS_PER_CALL = {"csm_ntxent_fwd": 2, "csm_decoder_assemble_bwd": 2, "csm_attention_bwd": 2, "csm_zero_async": 0}


class NativeError(RuntimeError):
    pass


def library_path():
    # CSMAE_LIB: a development build of the same library (e.g. the -DCSM_ATTN_TIMING variant of tools/attn_phase.py)
    return os.environ.get("CSMAE_LIB") or _build.LIB_PATH


def load():
    """Loads the shared library (never builds it implicitly on a GPU box: it must travel prebuilt)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise NativeError(
                f"{path} is missing: run `python __graft_entry__.py` (build()) first. "
                "csmae_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(path)
        lib.csm_last_error.restype = ctypes.c_char_p
        lib.csm_last_error.argtypes = []
        for name, argtypes in {**SIGNATURES, **DEV_SIGNATURES}.items():
            fn = getattr(lib, name)
            fn.restype = _c_int
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def _check(rc, name):
    if rc < 0:
        msg = load().csm_last_error().decode("utf-8", "replace")
        raise NativeError(f"{name} failed ({rc}): {msg}")
    return rc


def sm_count(device=None):
    dev = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    if dev not in _sm_count:
        _sm_count[dev] = _check(load().csm_device_check(dev), "csm_device_check")
    return _sm_count[dev]


_gemm_ws = {}


def enable_gemm_stream_k(device=None, enable=True):
    """Registers (or drops) the stream-K workspace of the forward / dgrad GEMMs on this device: zeroed device memory
    owned here.  Contract (include/csmae_b200.h): forward / dgrad GEMMs run on one stream at a time afterwards."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = load()
    if not enable:
        _check(lib.csm_gemm_set_workspace(0, 0), "csm_gemm_set_workspace")
        _gemm_ws.pop(dev.index or 0, None)
        return None
    key = dev.index or 0
    if key not in _gemm_ws:
        nbytes = _check(lib.csm_gemm_workspace_bytes(sm_count(dev)), "csm_gemm_workspace_bytes")
        _gemm_ws[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    ws = _gemm_ws[key]
    _check(lib.csm_gemm_set_workspace(ws.data_ptr(), ws.numel()), "csm_gemm_set_workspace")
    return ws


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def call(name, *args):
    """Raw call: tensors are converted to device pointers, the current stream is appended."""
    global launch_count
    lib = load()
    conv = [(_p(a) if (a is None or isinstance(a, torch.Tensor)) else a) for a in args]
    rc = getattr(lib, name)(*conv, _stream())
    launch_count += KERNELS_PER_CALL.get(name, 1)
    return _check(rc, name)
