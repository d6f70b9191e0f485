"""ctypes binding of libvrvq.so (include/vrvq.h).  PyTorch is used only for device memory and streams.

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
current, the calls raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvrvq.so")

CD = 8
ABI_VERSION = 1


class EncodeArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("B", C.c_int32), ("T", C.c_int32), ("input_dim", C.c_int32), ("n_codebooks", C.c_int32),
        ("codebook_size", C.c_int32), ("n_run", C.c_int32),
        ("blob", C.c_void_p),
        ("z", C.c_void_p), ("z_stride_b", C.c_int64), ("z_stride_d", C.c_int64),
        ("imp_map", C.c_void_p), ("imp_stride_b", C.c_int64),
        ("level_dev", C.c_void_p), ("level_stride", C.c_int64), ("level_host", C.c_float),
        ("codes", C.c_void_p), ("codes_stride_b", C.c_int64), ("codes_stride_q", C.c_int64),
        ("z_q", C.c_void_p), ("z_q_stride_b", C.c_int64), ("z_q_stride_d", C.c_int64),
        ("z_q_is", C.c_void_p), ("z_q_is_stride_b", C.c_int64), ("z_q_is_stride_q", C.c_int64), ("z_q_is_stride_d", C.c_int64),
        ("latents", C.c_void_p), ("latents_stride_b", C.c_int64), ("latents_stride_c", C.c_int64),
        ("mask", C.c_void_p), ("mask_stride_b", C.c_int64), ("mask_stride_q", C.c_int64),
        ("loss_pf", C.c_void_p), ("loss_pf_stride_b", C.c_int64), ("loss_pf_stride_q", C.c_int64),
        ("loss_masked_sum", C.c_void_p), ("kept", C.c_void_p),
    ]


class FromCodesArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("B", C.c_int32), ("T", C.c_int32), ("input_dim", C.c_int32), ("n_codebooks", C.c_int32),
        ("codebook_size", C.c_int32), ("n_run", C.c_int32),
        ("blob", C.c_void_p),
        ("codes", C.c_void_p), ("codes_stride_b", C.c_int64), ("codes_stride_q", C.c_int64),
        ("mask", C.c_void_p), ("mask_stride_b", C.c_int64), ("mask_stride_q", C.c_int64),
        ("z_q", C.c_void_p), ("z_q_stride_b", C.c_int64), ("z_q_stride_d", C.c_int64),
        ("z_p", C.c_void_p), ("z_p_stride_b", C.c_int64), ("z_p_stride_c", C.c_int64),
        ("z_q_is", C.c_void_p), ("z_q_is_stride_b", C.c_int64), ("z_q_is_stride_q", C.c_int64), ("z_q_is_stride_d", C.c_int64),
        ("error_flag", C.c_void_p),
    ]


# every symbol include/vrvq.h declares (tests check that the library exports all of them)
EXPORTS = [
    "vrvq_abi_version", "vrvq_last_error", "vrvq_supported", "vrvq_blob_bytes", "vrvq_pack_weights",
    "vrvq_blob_codebook", "vrvq_rvq_encode_f32", "vrvq_rvq_encode_launch_info", "vrvq_rvq_encode_kernel_name", "vrvq_from_codes_f32",
    "vrvq_search_latents_f32", "vrvq_generate_mask_hard_f32", "vrvq_mask_sum_f32", "vrvq_remask_f32",
    "vrvq_pack_codes_u16", "vrvq_unpack_codes_u16", "vrvq_conv3_packed_floats", "vrvq_pack_conv3_weights", "vrvq_snake_conv3_f32",
    "vrvq_conv3_tc_packed_floats", "vrvq_pack_conv3_tc_weights", "vrvq_snake_conv3_tc_f32", "vrvq_snake_f32",
    "vrvq_subnet_tail_usable", "vrvq_subnet_tail_f32", "vrvq_flat_tile_order",
]

_lib = None
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


class VrvqError(RuntimeError):
    pass


def lib():
    """Load libvrvq.so (building it first if only the sources are present and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build

            _build.build()
        except Exception as e:  # no silent fallback: the CUDA extension is the product
            raise ImportError(
                f"{LIB_PATH} is missing and could not be built ({e}); run `python -m vrvq_b200.build`. "
                "vrvq_b200 has no CPU or PyTorch fallback.") from e
    L = C.CDLL(LIB_PATH)
    L.vrvq_last_error.restype = C.c_char_p
    L.vrvq_blob_bytes.restype = C.c_size_t
    L.vrvq_blob_bytes.argtypes = [C.c_int] * 4
    L.vrvq_pack_weights.argtypes = [C.c_int] * 4 + [C.c_void_p] * 6 + [C.c_size_t]
    L.vrvq_blob_codebook.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    L.vrvq_rvq_encode_f32.argtypes = [C.POINTER(EncodeArgs), C.c_void_p]
    L.vrvq_rvq_encode_launch_info.argtypes = [C.POINTER(EncodeArgs), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.vrvq_rvq_encode_kernel_name.restype = C.c_char_p
    L.vrvq_rvq_encode_kernel_name.argtypes = [C.POINTER(EncodeArgs)]
    L.vrvq_from_codes_f32.argtypes = [C.POINTER(FromCodesArgs), C.c_void_p]
    L.vrvq_search_latents_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    L.vrvq_generate_mask_hard_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    L.vrvq_mask_sum_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.vrvq_remask_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_void_p]
    L.vrvq_pack_codes_u16.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vrvq_unpack_codes_u16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    L.vrvq_conv3_packed_floats.restype = C.c_size_t
    L.vrvq_conv3_packed_floats.argtypes = [C.c_int, C.c_int]
    L.vrvq_pack_conv3_weights.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.vrvq_snake_conv3_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    L.vrvq_conv3_tc_packed_floats.restype = C.c_size_t
    L.vrvq_conv3_tc_packed_floats.argtypes = [C.c_int, C.c_int]
    L.vrvq_pack_conv3_tc_weights.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.vrvq_snake_conv3_tc_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    L.vrvq_snake_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                 C.c_void_p]
    L.vrvq_subnet_tail_usable.argtypes = [C.c_int, C.c_int, C.c_int]
    L.vrvq_subnet_tail_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_void_p,
                                                                                                                  C.c_int64, C.c_void_p]
    if L.vrvq_abi_version() != ABI_VERSION:
        raise ImportError(f"libvrvq.so ABI {L.vrvq_abi_version()} != binding ABI {ABI_VERSION}; rebuild with python -m vrvq_b200.build")
    _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().vrvq_last_error().decode("utf-8", "replace")
        raise VrvqError(f"{what} failed (code {rc}): {msg}")


def current_stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda_f32(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise VrvqError(f"{name} is on {t.device}; vrvq_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise VrvqError(f"{name} must be float32, got {t.dtype}")


def count_launch(n=1):
    global launch_count
    launch_count += n
