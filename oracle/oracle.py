"""ctypes binding of the CPU oracle (oracle/ce_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of ce_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by codec_eval_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libce_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("ce_oracle.c", "ce_oracle_jpeg.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libce_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("valid", C.c_uint32),
        ("sse", C.c_uint64),
        ("dssim", C.c_double),
        ("ssimulacra2", C.c_double),
        ("butteraugli", C.c_double),
        ("psnr", C.c_double),
        ("butteraugli_pnorm3", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, f32p, f64p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_double)
        sz = C.c_size_t
        L.ceo_sse.restype = C.c_uint64
        L.ceo_sse.argtypes = [u8p, u8p, sz]
        L.ceo_psnr.restype = C.c_double
        L.ceo_psnr.argtypes = [u8p, u8p, sz, sz]
        L.ceo_psnr_from_sse.restype = C.c_double
        L.ceo_psnr_from_sse.argtypes = [C.c_uint64, sz, sz]
        L.ceo_xyb_roundtrip.argtypes = [u8p, sz, sz, u8p]
        L.ceo_xyb_roundtrip_libm.argtypes = [u8p, sz, sz, u8p]
        L.ceo_ssimulacra2.argtypes = [u8p, u8p, sz, sz, f64p]
        L.ceo_ssimulacra2_ex.argtypes = [u8p, u8p, sz, sz, f64p, f64p, C.POINTER(C.c_int)]
        L.ceo_ssimulacra2_scale0_planes.argtypes = [u8p, u8p, sz, sz, f32p]
        L.ceo_ssimulacra2_score_from_avgs.restype = C.c_double
        L.ceo_ssimulacra2_score_from_avgs.argtypes = [f64p, C.c_int]
        L.ceo_ssimulacra2_avgs_from_sums.argtypes = [f64p, sz, sz, f64p]
        L.ceo_dssim.argtypes = [u8p, u8p, sz, sz, f64p]
        L.ceo_dssim_ex.argtypes = [u8p, u8p, sz, sz, f64p, f64p, C.POINTER(C.c_int), f32p]
        L.ceo_dssim_rgbaf32.argtypes = [f32p, f32p, sz, sz, sz, f64p]
        L.ceo_dssim_from_scale_scores.restype = C.c_double
        L.ceo_dssim_from_scale_scores.argtypes = [f64p, C.c_int]
        L.ceo_rgb8_to_dssim_image.argtypes = [u8p, sz, sz, f32p]
        L.ceo_rgba8_to_dssim_image.argtypes = [u8p, sz, sz, f32p]
        L.ceo_butteraugli.argtypes = [u8p, u8p, sz, sz, C.c_float, f64p, f64p]
        L.ceo_butteraugli_ex.argtypes = [u8p, u8p, sz, sz, C.c_float, f64p, f64p, f32p]
        L.ceo_butteraugli_psycho.argtypes = [u8p, sz, sz, C.c_float, f32p]
        L.ceo_butteraugli_opsin.argtypes = [u8p, sz, sz, C.c_float, f32p]
        L.ceo_rgauss_blur.argtypes = [f32p, sz, sz, f32p]
        L.ceo_dssim_blur.argtypes = [f32p, sz, sz, f32p]
        L.ceo_ba_blur.argtypes = [f32p, sz, sz, C.c_float, f32p]
        L.ceo_ba_kernel.restype = C.c_int
        L.ceo_ba_kernel.argtypes = [C.c_float, f32p]
        L.ceo_srgb_lut.argtypes = [f32p]
        L.ceo_rgauss_coeffs.argtypes = [f32p]
        L.ceo_cbrtf.restype = C.c_float
        L.ceo_cbrtf.argtypes = [C.c_float]
        L.ceo_ba_fast_log2f.restype = C.c_float
        L.ceo_ba_fast_log2f.argtypes = [C.c_float]
        L.ceo_dssim_lab.argtypes = [C.c_float, C.c_float, C.c_float, f32p]
        L.ceo_evaluate_pair.argtypes = [u8p, u8p, sz, sz, C.c_uint32, C.c_float, C.POINTER(Result)]
        L.ceo_evaluate_batch.argtypes = [u8p, u8p, sz, sz, sz, C.c_uint32, C.c_float, C.c_int, C.POINTER(Result)]
        L.ceo_max_threads.restype = C.c_int
        L.ceo_jpeg_roundtrip.restype = C.c_int
        L.ceo_jpeg_roundtrip.argtypes = [u8p, sz, sz, C.c_int, C.c_int, u8p]
        L.ceo_jpeg_qtable.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint16)]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def sse(ref, dist) -> int:
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    assert r.size == d.size
    return int(lib().ceo_sse(rp, dp, r.size))


def psnr(ref, dist, w, h) -> float:
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    assert r.size == d.size == w * h * 3
    return float(lib().ceo_psnr(rp, dp, w, h))


def psnr_from_sse(s, w, h) -> float:
    return float(lib().ceo_psnr_from_sse(int(s), w, h))


def xyb_roundtrip(rgb, w, h) -> np.ndarray:
    r, rp = _u8(rgb)
    assert r.size == w * h * 3
    out = np.empty(w * h * 3, np.uint8)
    lib().ceo_xyb_roundtrip(rp, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def xyb_roundtrip_libm(rgb, w, h) -> np.ndarray:
    """xyb_roundtrip with this platform's cbrtf / powf (what Rust's f32::cbrt / f32::powf call on a glibc host)."""
    r, rp = _u8(rgb)
    assert r.size == w * h * 3
    out = np.empty(w * h * 3, np.uint8)
    lib().ceo_xyb_roundtrip_libm(rp, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def ssimulacra2(ref, dist, w, h) -> float:
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    out = C.c_double()
    st = lib().ceo_ssimulacra2(rp, dp, w, h, C.byref(out))
    if st:
        raise OracleError(st)
    return out.value


def ssimulacra2_ex(ref, dist, w, h):
    """-> (score, sums[nscales,18])"""
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    out = C.c_double()
    sums = np.zeros(6 * 18, np.float64)
    ns = C.c_int()
    st = lib().ceo_ssimulacra2_ex(rp, dp, w, h, C.byref(out), sums.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ns))
    if st:
        raise OracleError(st)
    return out.value, sums.reshape(6, 18)[: ns.value]


def ssimulacra2_scale0_planes(ref, dist, w, h) -> np.ndarray:
    """-> [3 channels, 7 (i1,i2,mu1,mu2,s11,s22,s12), h, w]"""
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    out = np.empty((3, 7, h, w), np.float32)
    lib().ceo_ssimulacra2_scale0_planes(rp, dp, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def dssim(ref, dist, w, h) -> float:
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    out = C.c_double()
    st = lib().ceo_dssim(rp, dp, w, h, C.byref(out))
    if st:
        raise OracleError(st)
    return out.value


def dssim_ex(ref, dist, w, h):
    """-> (dssim, per-scale scores, scale-0 ssim map)"""
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    out = C.c_double()
    sc = np.zeros(5, np.float64)
    ns = C.c_int()
    m = np.empty((h, w), np.float32)
    st = lib().ceo_dssim_ex(rp, dp, w, h, C.byref(out), sc.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ns),
                            m.ctypes.data_as(C.POINTER(C.c_float)))
    if st:
        raise OracleError(st)
    return out.value, sc[: ns.value], m


def dssim_rgbaf32(ref, dist, w, h, stride=None) -> float:
    r, rp = _f32(ref)
    d, dp = _f32(dist)
    out = C.c_double()
    st = lib().ceo_dssim_rgbaf32(rp, dp, w, h, stride or w, C.byref(out))
    if st:
        raise OracleError(st)
    return out.value


def rgb8_to_dssim_image(data, w, h) -> np.ndarray:
    r, rp = _u8(data)
    out = np.empty((h, w, 4), np.float32)
    lib().ceo_rgb8_to_dssim_image(rp, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def rgba8_to_dssim_image(data, w, h) -> np.ndarray:
    r, rp = _u8(data)
    out = np.empty((h, w, 4), np.float32)
    lib().ceo_rgba8_to_dssim_image(rp, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def butteraugli(ref, dist, w, h, intensity=80.0):
    """-> (max, pnorm3)"""
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    mx, pn = C.c_double(), C.c_double()
    st = lib().ceo_butteraugli(rp, dp, w, h, intensity, C.byref(mx), C.byref(pn))
    if st:
        raise OracleError(st)
    return mx.value, pn.value


def butteraugli_ex(ref, dist, w, h, intensity=80.0):
    """-> (max, pnorm3, diffmap[h,w])"""
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    mx, pn = C.c_double(), C.c_double()
    dm = np.empty((h, w), np.float32)
    st = lib().ceo_butteraugli_ex(rp, dp, w, h, intensity, C.byref(mx), C.byref(pn), dm.ctypes.data_as(C.POINTER(C.c_float)))
    if st:
        raise OracleError(st)
    return mx.value, pn.value, dm


def butteraugli_psycho(rgb, w, h, intensity=80.0) -> np.ndarray:
    """-> [10 (lf0..2, mf0..2, hf0..1, uhf0..1), h, w]"""
    r, rp = _u8(rgb)
    out = np.empty((10, h, w), np.float32)
    lib().ceo_butteraugli_psycho(rp, w, h, intensity, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def butteraugli_opsin(rgb, w, h, intensity=80.0) -> np.ndarray:
    r, rp = _u8(rgb)
    out = np.empty((3, h, w), np.float32)
    lib().ceo_butteraugli_opsin(rp, w, h, intensity, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def rgauss_blur(plane) -> np.ndarray:
    p, pp = _f32(plane)
    h, w = p.shape
    out = np.empty_like(p)
    lib().ceo_rgauss_blur(pp, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def dssim_blur(plane) -> np.ndarray:
    p, pp = _f32(plane)
    h, w = p.shape
    out = np.empty_like(p)
    lib().ceo_dssim_blur(pp, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def ba_blur(plane, sigma) -> np.ndarray:
    p, pp = _f32(plane)
    h, w = p.shape
    out = np.empty_like(p)
    lib().ceo_ba_blur(pp, w, h, sigma, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def ba_kernel(sigma):
    buf = np.zeros(65, np.float32)
    r = lib().ceo_ba_kernel(sigma, buf.ctypes.data_as(C.POINTER(C.c_float)))
    return r, buf[: 2 * r + 1].copy()


def srgb_lut() -> np.ndarray:
    out = np.empty(256, np.float32)
    lib().ceo_srgb_lut(out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def rgauss_coeffs() -> np.ndarray:
    out = np.empty(9, np.float32)
    lib().ceo_rgauss_coeffs(out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


FLAG_DSSIM, FLAG_SSIM2, FLAG_BUTTERAUGLI, FLAG_PSNR, FLAG_XYB = 1, 2, 4, 8, 16


def evaluate_pair(ref, dist, w, h, flags, intensity=80.0) -> Result:
    r, rp = _u8(ref)
    d, dp = _u8(dist)
    out = Result()
    lib().ceo_evaluate_pair(rp, dp, w, h, flags, intensity, C.byref(out))
    return out


def evaluate_batch(refs, dists, w, h, flags, intensity=80.0, threads=0):
    """refs, dists: uint8 arrays [n, h, w, 3]; OpenMP over pairs."""
    r, rp = _u8(refs)
    d, dp = _u8(dists)
    n = r.size // (w * h * 3)
    out = (Result * n)()
    lib().ceo_evaluate_batch(rp, dp, n, w, h, flags, intensity, threads, out)
    return list(out)


def max_threads() -> int:
    return int(lib().ceo_max_threads())


def jpeg_roundtrip(rgb, w, h, quality, subsampling=2) -> np.ndarray:
    """Baseline JPEG encode -> decode in the sample domain (ce_oracle_jpeg.c); uint8 [h, w, 3]."""
    r, rp = _u8(rgb)
    assert r.size == w * h * 3
    out = np.empty((h, w, 3), np.uint8)
    st = lib().ceo_jpeg_roundtrip(rp, w, h, int(quality), int(subsampling), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if st != 0:
        raise OracleError(st)
    return out


def jpeg_qtable(chroma: bool, quality: int) -> np.ndarray:
    t = np.zeros(64, np.uint16)
    lib().ceo_jpeg_qtable(int(chroma), int(quality), t.ctypes.data_as(C.POINTER(C.c_uint16)))
    return t
