"""ctypes binding of libce_gpu.so (include/ce_gpu.h).  Fails loudly if the library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CE_LIB_PATH: an experiment build made with `python -m codec_eval_b200.build --out <path> -D...` (A/B runs only)
LIB_PATH = os.environ.get("CE_LIB_PATH") or os.path.join(_HERE, "libce_gpu.so")
_EXPERIMENT = bool(os.environ.get("CE_LIB_PATH"))

CE_OK = 0
CE_ERR_DIMENSION_MISMATCH = 1
CE_ERR_METRIC_CALCULATION = 2
CE_ERR_INVALID_ARGUMENT = 3
CE_ERR_CUDA = 4
CE_ERR_OUT_OF_MEMORY = 5

VALID_DSSIM, VALID_SSIMULACRA2, VALID_BUTTERAUGLI, VALID_PSNR = 1, 2, 4, 8


class CeMetricConfig(C.Structure):
    _fields_ = [("dssim", C.c_uint8), ("ssimulacra2", C.c_uint8), ("butteraugli", C.c_uint8), ("psnr", C.c_uint8),
                ("xyb_roundtrip", C.c_uint8)]


class CePair(C.Structure):
    _fields_ = [("ref", C.c_void_p), ("dist", C.c_void_p), ("ref_len", C.c_size_t), ("dist_len", C.c_size_t),
                ("width", C.c_uint32), ("height", C.c_uint32), ("ref_id", C.c_uint32), ("reserved", C.c_uint32)]


class CeResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("valid", C.c_uint32), ("sse", C.c_uint64), ("dssim", C.c_double),
                ("ssimulacra2", C.c_double), ("butteraugli", C.c_double), ("psnr", C.c_double),
                ("butteraugli_pnorm3", C.c_double)]


# every symbol include/ce_gpu.h declares (tests check that the library exports all of them)
EXPORTS = [
    "ce_ctx_create", "ce_ctx_destroy", "ce_ctx_set_stream", "ce_last_error", "ce_launch_count", "ce_version",
    "ce_profile_enable", "ce_profile_reset", "ce_profile_report",
    "ce_host_register", "ce_host_unregister", "ce_host_alloc", "ce_host_free",
    "ce_sub_batch_capacity", "ce_evaluate_batch", "ce_evaluate_batch_device", "ce_evaluate_batch_device_grouped", "ce_psnr", "ce_ssimulacra2", "ce_butteraugli", "ce_dssim_rgb8",
    "ce_dssim_rgbaf32", "ce_rgb8_to_dssim_image", "ce_rgba8_to_dssim_image", "ce_xyb_roundtrip",
    "ce_reference_create", "ce_reference_compare", "ce_reference_compare_many", "ce_reference_destroy",
    "ce_jpeg_roundtrip", "ce_jpeg_roundtrip_device", "ce_evaluate_jpeg_sweep", "ce_transform_to_srgb",
    "ce_debug_ssim2_sums", "ce_debug_ssim2_scale0_planes", "ce_debug_dssim_scales", "ce_debug_butteraugli_diffmap",
    "ce_debug_butteraugli_psycho", "ce_debug_butteraugli_opsin", "ce_debug_ba_blur",
]

_lib = None


def source_hash() -> str:
    """sha256/16 over csrc/* and include/ce_gpu.h (names and contents, sorted) -- what build.py stamps into
    ce_version().  Empty string when the sources are not beside the package (a binary-only install)."""
    import hashlib

    csrc = os.path.join(_HERE, "csrc")
    header = os.path.join(_HERE, "..", "include", "ce_gpu.h")
    if not os.path.isdir(csrc) or not os.path.exists(header):
        return ""
    hh = hashlib.sha256()
    for path in sorted(os.path.join(csrc, f) for f in os.listdir(csrc)) + [header]:
        hh.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as f:
            hh.update(f.read())
    return hh.hexdigest()[:16]


def _check_source_hash(version: str):
    """A library built from other sources than the ones beside it must not be taken for them."""
    want = source_hash()
    if not want or os.environ.get("CE_ALLOW_STALE_LIB") == "1":
        return
    if _EXPERIMENT:   # same sources, other -D flags: "src:<hash>+<flags>"
        version = version.split("+", 1)[0]
    got = version.rsplit("src:", 1)[-1] if "src:" in version else "unhashed"
    if got != want:
        raise RuntimeError(f"{LIB_PATH} was built from sources with hash {got}, the tree has {want}: rebuild with "
                           "`python -m codec_eval_b200.build` (or __graft_entry__.build())")


def load():
    """Load libce_gpu.so and declare prototypes.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _EXPERIMENT and os.environ.get("CE_ALLOW_STALE_LIB") != "1" and os.path.isdir(os.path.join(_HERE, "csrc")):
        # the library must be the build of the sources beside it: rebuild when it is missing or stale (nvcc is part of
        # the image on both the CPU and the GPU box); without nvcc this raises -- there is no CPU fallback
        from . import build as _build

        try:
            _build.ensure_built()
        except Exception as e:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing and could not be built: {e}. codec_eval_b200 has no CPU fallback.")
            # a stale library is refused by _check_source_hash below
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m codec_eval_b200.build` "
            "(nvcc, sm_100a). codec_eval_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ce_version.restype = C.c_char_p
    _check_source_hash(L.ce_version().decode())
    vp, sz, u8p = C.c_void_p, C.c_size_t, C.c_void_p
    f64p, f32p = C.POINTER(C.c_double), C.c_void_p
    cfgp, resp = C.POINTER(CeMetricConfig), C.POINTER(CeResult)
    L.ce_ctx_create.argtypes = [C.POINTER(vp), C.c_int, sz]
    L.ce_ctx_destroy.argtypes = [vp]
    L.ce_ctx_destroy.restype = None
    L.ce_ctx_set_stream.argtypes = [vp, vp]
    L.ce_last_error.argtypes = [vp]
    L.ce_last_error.restype = C.c_char_p
    L.ce_launch_count.argtypes = [vp]
    L.ce_launch_count.restype = C.c_uint64
    L.ce_version.restype = C.c_char_p
    L.ce_profile_enable.argtypes = [vp, C.c_int]
    L.ce_profile_reset.argtypes = [vp]
    L.ce_profile_report.argtypes = [vp, C.c_char_p, sz]
    L.ce_profile_report.restype = sz
    L.ce_host_register.argtypes = [vp, vp, sz]
    L.ce_host_unregister.argtypes = [vp, vp]
    L.ce_host_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.ce_host_free.argtypes = [vp, vp]
    L.ce_host_free.restype = None
    L.ce_sub_batch_capacity.argtypes = [vp, cfgp, C.c_uint32, C.c_uint32, C.POINTER(sz)]
    L.ce_evaluate_batch.argtypes = [vp, C.POINTER(CePair), sz, cfgp, C.c_float, resp]
    L.ce_evaluate_batch_device.argtypes = [vp, vp, vp, sz, C.c_uint32, C.c_uint32, cfgp, C.c_float, resp]
    L.ce_evaluate_batch_device_grouped.argtypes = [vp, vp, sz, vp, sz, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, cfgp,
                                                   C.c_float, resp]
    L.ce_psnr.argtypes = [vp, u8p, sz, u8p, sz, sz, sz, f64p, C.POINTER(C.c_uint64)]
    L.ce_ssimulacra2.argtypes = [vp, u8p, sz, u8p, sz, sz, sz, f64p]
    L.ce_butteraugli.argtypes = [vp, u8p, sz, u8p, sz, sz, sz, C.c_float, f64p, f64p]
    L.ce_dssim_rgb8.argtypes = [vp, u8p, sz, u8p, sz, sz, sz, f64p]
    L.ce_dssim_rgbaf32.argtypes = [vp, f32p, sz, sz, sz, f32p, sz, sz, sz, f64p]
    L.ce_rgb8_to_dssim_image.argtypes = [vp, u8p, sz, sz, sz, f32p]
    L.ce_rgba8_to_dssim_image.argtypes = [vp, u8p, sz, sz, sz, f32p]
    L.ce_xyb_roundtrip.argtypes = [vp, u8p, sz, sz, sz, u8p]
    L.ce_reference_create.argtypes = [vp, u8p, sz, sz, sz, cfgp, C.POINTER(vp)]
    L.ce_reference_compare.argtypes = [vp, vp, u8p, sz, C.c_float, resp]
    L.ce_reference_compare_many.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(sz), sz, C.c_float, resp]
    L.ce_reference_destroy.argtypes = [vp]
    L.ce_reference_destroy.restype = None
    L.ce_jpeg_roundtrip.argtypes = [vp, u8p, sz, sz, sz, C.c_int, C.c_int, u8p]
    L.ce_jpeg_roundtrip_device.argtypes = [vp, vp, sz, C.c_uint32, C.c_uint32, C.POINTER(C.c_int), sz, C.c_int, vp]
    L.ce_evaluate_jpeg_sweep.argtypes = [vp, C.POINTER(vp), sz, C.c_uint32, C.c_uint32, C.POINTER(C.c_int), sz, C.c_int, cfgp,
                                         C.c_float, resp]
    L.ce_transform_to_srgb.argtypes = [vp, u8p, sz, sz, sz, u8p, sz, u8p]
    L.ce_debug_ssim2_sums.argtypes = [vp, u8p, u8p, sz, sz, f64p, C.POINTER(C.c_int)]
    L.ce_debug_ssim2_scale0_planes.argtypes = [vp, u8p, u8p, sz, sz, f32p]
    L.ce_debug_dssim_scales.argtypes = [vp, u8p, u8p, sz, sz, f64p, C.POINTER(C.c_int), f32p]
    L.ce_debug_butteraugli_diffmap.argtypes = [vp, u8p, u8p, sz, sz, C.c_float, f32p]
    L.ce_debug_butteraugli_psycho.argtypes = [vp, u8p, sz, sz, C.c_float, f32p]
    L.ce_debug_butteraugli_opsin.argtypes = [vp, u8p, sz, sz, C.c_float, f32p]
    L.ce_debug_ba_blur.argtypes = [vp, f32p, sz, sz, C.c_float, f32p]
    _lib = L
    return L
