"""ctypes binding of libmst_b200.so (the C ABI declared in include/mst_b200.h).

There is NO CPU fallback: importing the package works without the library (so CPU-only tooling can import the module
surface), but every compute call goes through `lib()` which raises if the .so is missing, and every op raises if its
tensors are not CUDA tensors.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libmst_b200.so")

MST_MAX_ENC_BLOCKS = 16
FX_NPARAMS = 20
FX_EQ, FX_COMP, FX_IMAGER, FX_GAIN, FX_RMSNORM = 1, 2, 4, 8, 16
FX_ALL = FX_EQ | FX_COMP | FX_IMAGER | FX_GAIN | FX_RMSNORM
TCN_F16F8, TCN_BF16X3 = 0, 1          # MST_TCN_F16F8 / MST_TCN_BF16X3
TCN_F16F8_RANGE = 448.0               # MST_TCN_F16F8_RANGE


class EncConfig(ctypes.Structure):
    _fields_ = [("n_blocks", c_int),
                ("channels", c_int * (MST_MAX_ENC_BLOCKS + 1)),
                ("kernels", c_int * MST_MAX_ENC_BLOCKS),
                ("strides", c_int * MST_MAX_ENC_BLOCKS)]


class TcnConfig(ctypes.Structure):
    _fields_ = [("n_blocks", c_int), ("n_inputs", c_int), ("n_outputs", c_int), ("channels", c_int),
                ("kernel_size", c_int), ("dilation_growth", c_int), ("stack_size", c_int), ("cond_dim", c_int)]


# name -> (restype, argtypes); must list EVERY symbol include/mst_b200.h declares (tests/test_cabi_symbols.py)
SIGNATURES = {
    "mst_last_error": (c_char_p, []),
    "mst_version": (c_int, []),
    "mst_device_check": (c_int, [c_int]),
    "mst_conv1d_fold_bn": (c_int, [c_void_p] * 6 + [c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mst_enc_conv1d": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    "mst_enc_mean_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mst_rows_reduce": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "mst_enc_packed_bytes": (c_size_t, [POINTER(EncConfig)]),
    "mst_enc_pack": (c_int, [POINTER(EncConfig), POINTER(c_void_p), c_void_p, c_void_p]),
    "mst_enc_workspace_bytes": (c_size_t, [POINTER(EncConfig), c_int, c_int]),
    "mst_enc_forward": (c_int, [POINTER(EncConfig), c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "mst_tcn_packed_bytes": (c_size_t, [POINTER(TcnConfig)]),
    "mst_tcn_pack": (c_int, [POINTER(TcnConfig), POINTER(c_void_p), c_void_p, c_void_p]),
    "mst_tcn_film_precompute": (c_int, [POINTER(TcnConfig), c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mst_tcn_workspace_bytes": (c_size_t, [POINTER(TcnConfig), c_int, c_int]),
    "mst_tcn_forward": (c_int, [POINTER(TcnConfig), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "mst_tcn_block0_forward": (c_int, [POINTER(TcnConfig), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                       c_int, c_void_p, c_void_p]),
    "mst_tcn_layer_forward": (c_int, [POINTER(TcnConfig), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "mst_tcn_block_forward": (c_int, [POINTER(TcnConfig), c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                      c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "mst_fx_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mst_fx_chain_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_size_t,
                                     c_void_p]),
    "mst_biquad_cascade": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mst_stereo_stats": (c_int, [c_void_p, c_int, ctypes.c_longlong, c_void_p, c_void_p]),
    "mst_stereo_mix": (c_int, [c_void_p, c_void_p, c_void_p, c_int, ctypes.c_longlong, c_void_p]),
    "mst_block_energy": (c_int, [c_void_p, c_int, ctypes.c_longlong, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mst_haas": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mst_row_absmax": (c_int, [c_void_p, c_int, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p]),
    "mst_stft_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mst_stft_mag_mean": (c_int, [c_void_p, c_int, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "mst_fir_filtfilt_workspace_bytes": (c_size_t, [c_int, ctypes.c_longlong, c_int]),
    "mst_fir_filtfilt": (c_int, [c_void_p, c_int, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_int, c_void_p, c_void_p,
                                 ctypes.c_longlong, c_void_p, c_size_t, c_void_p]),
    "mst_fft_convolve_workspace_bytes": (c_size_t, [ctypes.c_longlong, ctypes.c_longlong]),
    "mst_fft_convolve": (c_int, [c_void_p, ctypes.c_longlong, ctypes.c_longlong, c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                                 c_int, ctypes.c_longlong, c_float, c_float, c_void_p, ctypes.c_longlong, c_void_p, c_size_t,
                                 c_void_p]),
    "mst_algo_reverb_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mst_algo_reverb": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mst_pcm_decode": (c_int, [c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, ctypes.c_longlong, c_void_p]),
    "mst_pcm_encode_mix": (c_int, [c_void_p, c_int, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library.  Raises (loudly) if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m music_mixing_style_transfer_b200.build` "
                "(or __graft_entry__.build()). This engine has no CPU / PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().mst_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libmst_b200 {what}: {last_error()}")


def ptr(t, rows_strided: bool = False) -> int:
    """Raw device pointer of a CUDA tensor (None -> NULL).  The C ABI takes dense buffers: a tensor whose strides are not the
    contiguous ones is refused instead of being read with the wrong layout (`rows_strided`: a 2-D view whose rows are dense but
    may lie further apart than their length, for the entry points that take a row stride)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libmst_b200 op called with a CPU tensor: this engine has no CPU fallback")
    if rows_strided:
        if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1) or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
            raise RuntimeError(f"libmst_b200 op needs dense rows, got strides {tuple(t.stride())} for shape {tuple(t.shape)}")
    elif not t.is_contiguous():
        raise RuntimeError(f"libmst_b200 op called with a non-contiguous tensor (shape {tuple(t.shape)}, strides {tuple(t.stride())}); "
                           "call .contiguous() first")
    return t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda_f32(t, name: str):
    import torch

    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the B200 engine only runs on CUDA tensors (no CPU fallback); "
                           "move the module and its inputs to a CUDA device")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()
