"""Host-side mirror of codec-eval's `src/metrics` API over libce_gpu (the C ABI in include/ce_gpu.h).

Same names, argument meaning and error behaviour as the reference:
  MetricConfig / MetricResult / PerceptionLevel   src/metrics/mod.rs:45-296
  calculate_psnr                                  src/metrics/mod.rs:312-331
  calculate_ssimulacra2                           src/metrics/ssimulacra2.rs:59-100
  calculate_butteraugli[_with_intensity]          src/metrics/butteraugli.rs:45-136
  calculate_dssim / rgb8_to_dssim_image / rgba8_  src/metrics/dssim.rs:40-143
  xyb_roundtrip                                   src/metrics/xyb.rs:225-253
plus the batched GPU entry (`GpuMetrics.evaluate_batch`) that EvalSession::calculate_metrics
dispatches into (src/eval/session.rs:437-497).  All arithmetic runs in CUDA kernels; this file
only marshals buffers.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
import weakref
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------- errors (src/error.rs)
class Error(Exception):
    """codec_eval::Error"""


class DimensionMismatch(Error):  # src/error.rs:33
    def __init__(self, expected: Tuple[int, int], actual: Tuple[int, int]):
        super().__init__(f"Dimension mismatch: expected {expected[0]}x{expected[1]}, got {actual[0]}x{actual[1]}")
        self.expected, self.actual = expected, actual


class MetricCalculation(Error):  # src/error.rs:42
    def __init__(self, metric: str, reason: str):
        super().__init__(f"Failed to calculate {metric}: {reason}")
        self.metric, self.reason = metric, reason


class CudaError(Error):
    pass


# ----------------------------------------------------------------------------- config / result types
@dataclass
class MetricConfig:  # src/metrics/mod.rs:45-63
    dssim: bool = False
    ssimulacra2: bool = False
    butteraugli: bool = False
    psnr: bool = False
    xyb_roundtrip: bool = False

    @staticmethod
    def all() -> "MetricConfig":  # mod.rs:67-76
        return MetricConfig(True, True, True, True, False)

    @staticmethod
    def fast() -> "MetricConfig":  # mod.rs:79-88
        return MetricConfig(False, False, False, True, False)

    @staticmethod
    def perceptual() -> "MetricConfig":  # mod.rs:91-100
        return MetricConfig(True, True, True, False, False)

    @staticmethod
    def perceptual_xyb() -> "MetricConfig":  # mod.rs:108-117
        return MetricConfig(True, True, True, False, True)

    @staticmethod
    def ssimulacra2_only() -> "MetricConfig":  # mod.rs:120-129
        return MetricConfig(False, True, False, False, False)

    def with_xyb_roundtrip(self) -> "MetricConfig":  # mod.rs:132-136
        return MetricConfig(self.dssim, self.ssimulacra2, self.butteraugli, self.psnr, True)

    def _c(self) -> _lib.CeMetricConfig:
        return _lib.CeMetricConfig(int(self.dssim), int(self.ssimulacra2), int(self.butteraugli), int(self.psnr),
                                   int(self.xyb_roundtrip))


class PerceptionLevel(enum.Enum):  # src/metrics/mod.rs:172-296
    Imperceptible = "IMP"
    Marginal = "MAR"
    Subtle = "SUB"
    Noticeable = "NOT"
    Degraded = "DEG"

    @staticmethod
    def from_dssim(d: float) -> "PerceptionLevel":
        if d < 0.0003:
            return PerceptionLevel.Imperceptible
        if d < 0.0007:
            return PerceptionLevel.Marginal
        if d < 0.0015:
            return PerceptionLevel.Subtle
        if d < 0.003:
            return PerceptionLevel.Noticeable
        return PerceptionLevel.Degraded

    @staticmethod
    def from_ssimulacra2(s: float) -> "PerceptionLevel":
        if s > 90.0:
            return PerceptionLevel.Imperceptible
        if s > 80.0:
            return PerceptionLevel.Marginal
        if s > 70.0:
            return PerceptionLevel.Subtle
        if s > 50.0:
            return PerceptionLevel.Noticeable
        return PerceptionLevel.Degraded

    @staticmethod
    def from_butteraugli(s: float) -> "PerceptionLevel":
        if s < 1.0:
            return PerceptionLevel.Imperceptible
        if s < 2.0:
            return PerceptionLevel.Marginal
        if s < 3.0:
            return PerceptionLevel.Subtle
        if s < 5.0:
            return PerceptionLevel.Noticeable
        return PerceptionLevel.Degraded

    def max_dssim(self) -> float:
        return {"IMP": 0.0003, "MAR": 0.0007, "SUB": 0.0015, "NOT": 0.003, "DEG": math.inf}[self.value]

    def min_ssimulacra2(self) -> float:
        return {"IMP": 90.0, "MAR": 80.0, "SUB": 70.0, "NOT": 50.0, "DEG": -math.inf}[self.value]

    def max_butteraugli(self) -> float:
        return {"IMP": 1.0, "MAR": 2.0, "SUB": 3.0, "NOT": 5.0, "DEG": math.inf}[self.value]

    def code(self) -> str:
        return self.value

    def __str__(self) -> str:
        return self.name


@dataclass
class MetricResult:  # src/metrics/mod.rs:139-149 (+ raw SSE and the 3-norm the GPU path also returns)
    dssim: Optional[float] = None
    ssimulacra2: Optional[float] = None
    butteraugli: Optional[float] = None
    psnr: Optional[float] = None
    sse: Optional[int] = None
    butteraugli_pnorm3: Optional[float] = None

    def perception_level(self) -> Optional[PerceptionLevel]:
        return None if self.dssim is None else PerceptionLevel.from_dssim(self.dssim)

    def perception_level_ssimulacra2(self) -> Optional[PerceptionLevel]:
        return None if self.ssimulacra2 is None else PerceptionLevel.from_ssimulacra2(self.ssimulacra2)

    def perception_level_butteraugli(self) -> Optional[PerceptionLevel]:
        return None if self.butteraugli is None else PerceptionLevel.from_butteraugli(self.butteraugli)


class ColorProfile:
    """src/metrics/icc.rs:31-56: `ColorProfile::Srgb` (no transform) or `ColorProfile::Icc(bytes)`."""

    __slots__ = ("icc",)

    def __init__(self, icc: Optional[bytes] = None):
        self.icc = bytes(icc) if icc else None

    Srgb: "ColorProfile"  # set below

    @staticmethod
    def Icc(data: bytes) -> "ColorProfile":
        p = ColorProfile()
        p.icc = bytes(data)  # kept even if empty, like the Rust variant
        return p

    def is_srgb(self) -> bool:  # icc.rs:43-45
        return self.icc is None

    @staticmethod
    def from_icc_bytes(icc: Optional[bytes]) -> "ColorProfile":  # icc.rs:49-54: None / empty -> Srgb
        return ColorProfile.Icc(icc) if icc else ColorProfile()

    def __repr__(self) -> str:
        return "Srgb" if self.icc is None else f"Icc({len(self.icc)} bytes)"


ColorProfile.Srgb = ColorProfile()


def _result_from_c(r: _lib.CeResult) -> MetricResult:
    v = r.valid
    return MetricResult(
        dssim=r.dssim if v & _lib.VALID_DSSIM else None,
        ssimulacra2=r.ssimulacra2 if v & _lib.VALID_SSIMULACRA2 else None,
        butteraugli=r.butteraugli if v & _lib.VALID_BUTTERAUGLI else None,
        psnr=r.psnr if v & _lib.VALID_PSNR else None,
        sse=int(r.sse) if v & _lib.VALID_PSNR else None,
        butteraugli_pnorm3=r.butteraugli_pnorm3 if v & _lib.VALID_BUTTERAUGLI else None,
    )


def _as_u8(buf) -> np.ndarray:
    a = np.asarray(buf)
    if a.dtype != np.uint8:
        raise TypeError("expected uint8 pixel data")
    return np.ascontiguousarray(a).reshape(-1)


# ----------------------------------------------------------------------------- the context
class GpuMetrics:
    """One CUDA device + stream + workspace (ce_ctx).  Single owner, like GpuSsim2
    (crates/codec-iter/src/gpu.rs:21-38)."""

    def __init__(self, device: int = 0, workspace_bytes: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        st = self._L.ce_ctx_create(C.byref(h), device, workspace_bytes)
        if st != _lib.CE_OK:
            raise CudaError(f"ce_ctx_create failed ({st}): {self._L.ce_last_error(None).decode()}")
        self._h = h
        self.device = device
        self._refs = weakref.WeakSet()   # live GpuReference handles: closed with the context

    def close(self):
        if getattr(self, "_h", None):
            for r in list(getattr(self, "_refs", ())):
                r.close()
            self._L.ce_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing
    def set_stream(self, cuda_stream: int):
        self._L.ce_ctx_set_stream(self._h, C.c_void_p(cuda_stream))

    def last_error(self) -> str:
        return self._L.ce_last_error(self._h).decode()

    def launch_count(self) -> int:
        return int(self._L.ce_launch_count(self._h))

    def profile(self, enable: bool = True, reset: bool = True):
        """Per-kernel CUDA-event timing (name -> launches, ms, algorithmic bytes)."""
        self._L.ce_profile_enable(self._h, int(enable))
        if reset:
            self._L.ce_profile_reset(self._h)

    def profile_report(self):
        n = self._L.ce_profile_report(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self._L.ce_profile_report(self._h, buf, n + 1)
        rows = {}
        for line in buf.value.decode().splitlines():
            name, launches, ms, nbytes, nbytes_pp = line.split("\t")
            rows[name] = {"launches": int(launches), "ms": float(ms), "bytes": float(nbytes), "bytes_per_pair": float(nbytes_pp)}
        return rows

    def _raise(self, st: int, metric: str, expected=None, actual=None):
        if st == _lib.CE_OK:
            return
        if st == _lib.CE_ERR_DIMENSION_MISMATCH:
            raise DimensionMismatch(expected or (0, 0), actual or (0, 0))
        if st == _lib.CE_ERR_METRIC_CALCULATION:
            raise MetricCalculation(metric, self.last_error())
        if st == _lib.CE_ERR_INVALID_ARGUMENT:
            raise AssertionError(self.last_error() or "invalid argument")
        raise CudaError(f"{metric}: status {st}: {self.last_error()}")

    # -- single-pair mirrors
    def calculate_psnr(self, reference, test, width: int, height: int) -> float:
        """src/metrics/mod.rs:312-331; length mismatches assert, like the reference."""
        r, t = _as_u8(reference), _as_u8(test)
        assert r.size == t.size, "assertion failed: reference.len() == test.len()"
        assert r.size == width * height * 3, "assertion failed: reference.len() == width * height * 3"
        out = C.c_double()
        sse = C.c_uint64()
        st = self._L.ce_psnr(self._h, r.ctypes.data, r.size, t.ctypes.data, t.size, width, height, C.byref(out), C.byref(sse))
        self._raise(st, "PSNR")
        return out.value

    def calculate_sse(self, reference, test, width: int, height: int) -> int:
        r, t = _as_u8(reference), _as_u8(test)
        out = C.c_double()
        sse = C.c_uint64()
        st = self._L.ce_psnr(self._h, r.ctypes.data, r.size, t.ctypes.data, t.size, width, height, C.byref(out), C.byref(sse))
        self._raise(st, "PSNR")
        return int(sse.value)

    def calculate_ssimulacra2(self, reference, test, width: int, height: int) -> float:
        r, t = _as_u8(reference), _as_u8(test)
        out = C.c_double()
        st = self._L.ce_ssimulacra2(self._h, r.ctypes.data, r.size, t.ctypes.data, t.size, width, height, C.byref(out))
        self._raise(st, "SSIMULACRA2", (width, height), (t.size // 3 // max(height, 1), height))
        return out.value

    def calculate_butteraugli_with_intensity(self, reference, test, width: int, height: int, intensity_target: float,
                                             return_pnorm: bool = False):
        r, t = _as_u8(reference), _as_u8(test)
        out, pn = C.c_double(), C.c_double()
        st = self._L.ce_butteraugli(self._h, r.ctypes.data, r.size, t.ctypes.data, t.size, width, height,
                                    intensity_target, C.byref(out), C.byref(pn))
        self._raise(st, "Butteraugli", (width, height), (t.size // 3 // max(height, 1), height))
        return (out.value, pn.value) if return_pnorm else out.value

    def calculate_butteraugli(self, reference, test, width: int, height: int, return_pnorm: bool = False):
        return self.calculate_butteraugli_with_intensity(reference, test, width, height, 80.0, return_pnorm)

    def calculate_dssim_rgb8(self, reference, test, width: int, height: int) -> float:
        """rgb8_to_dssim_image x2 + calculate_dssim fused (src/eval/session.rs:467-476)."""
        r, t = _as_u8(reference), _as_u8(test)
        out = C.c_double()
        st = self._L.ce_dssim_rgb8(self._h, r.ctypes.data, r.size, t.ctypes.data, t.size, width, height, C.byref(out))
        self._raise(st, "DSSIM", (width, height), (t.size // 3 // max(height, 1), height))
        return out.value

    def calculate_dssim(self, reference: np.ndarray, test: np.ndarray, viewing=None) -> float:
        """src/metrics/dssim.rs:40-71: linear RGBA f32 images [h, w, 4]; `viewing` is ignored (dssim.rs:43)."""
        ra = np.ascontiguousarray(reference, dtype=np.float32)
        ta = np.ascontiguousarray(test, dtype=np.float32)
        if ra.ndim != 3 or ta.ndim != 3 or ra.shape[2] != 4 or ta.shape[2] != 4:
            raise TypeError("expected [h, w, 4] float32 images")
        out = C.c_double()
        st = self._L.ce_dssim_rgbaf32(self._h, ra.ctypes.data, ra.shape[1], ra.shape[0], ra.shape[1], ta.ctypes.data,
                                      ta.shape[1], ta.shape[0], ta.shape[1], C.byref(out))
        self._raise(st, "DSSIM", (ra.shape[1], ra.shape[0]), (ta.shape[1], ta.shape[0]))
        return out.value

    def rgb8_to_dssim_image(self, data, width: int, height: int) -> np.ndarray:
        d = _as_u8(data)
        out = np.empty((height, width, 4), np.float32)
        st = self._L.ce_rgb8_to_dssim_image(self._h, d.ctypes.data, d.size, width, height, out.ctypes.data)
        self._raise(st, "DSSIM")
        return out

    def rgba8_to_dssim_image(self, data, width: int, height: int) -> np.ndarray:
        d = _as_u8(data)
        out = np.empty((height, width, 4), np.float32)
        st = self._L.ce_rgba8_to_dssim_image(self._h, d.ctypes.data, d.size, width, height, out.ctypes.data)
        self._raise(st, "DSSIM")
        return out

    def xyb_roundtrip(self, rgb, width: int, height: int) -> np.ndarray:
        d = _as_u8(rgb)
        assert d.size == width * height * 3, "Buffer size mismatch"
        out = np.empty(d.size, np.uint8)
        st = self._L.ce_xyb_roundtrip(self._h, d.ctypes.data, d.size, width, height, out.ctypes.data)
        self._raise(st, "XYB")
        return out

    # -- batched entries
    def evaluate_batch_raw(self, pairs: Sequence[Tuple[np.ndarray, np.ndarray, int, int]], config: MetricConfig,
                           intensity_target: float = 80.0):
        """pairs: (reference_u8, test_u8, width, height).  Returns the ce_result array (status per pair)."""
        n = len(pairs)
        arr = (_lib.CePair * max(n, 1))()
        keep = []
        ref_ids = {}   # pairs that hand over the same reference buffer share a ref_id (= one upload, one pre-processing)
        for i, (r, t, w, h) in enumerate(pairs):
            r, t = _as_u8(r), _as_u8(t)
            keep.append((r, t))
            rid = ref_ids.setdefault(r.ctypes.data, len(ref_ids))
            arr[i] = _lib.CePair(r.ctypes.data, t.ctypes.data, r.size, t.size, w, h, rid, 0)
        return self.evaluate_pair_table(arr, n, config, intensity_target)

    def evaluate_pair_table(self, table, n: int, config: MetricConfig, intensity_target: float = 80.0):
        """A prebuilt ce_pair[n] table (host pointers the caller keeps alive) -> ce_result[n]."""
        out = (_lib.CeResult * max(n, 1))()
        cfg = config._c()
        st = self._L.ce_evaluate_batch(self._h, table, n, C.byref(cfg), intensity_target, out)
        if st != _lib.CE_OK:
            self._raise(st, "batch")
        return out

    @staticmethod
    def _failing_metric(config: MetricConfig, width: int, height: int) -> str:
        """Name for Error::MetricCalculation.metric of a failed pair: the first enabled metric, in calculate_metrics'
        order (src/eval/session.rs:437-497), that can fail for these dimensions."""
        small = width < 8 or height < 8
        if small and config.ssimulacra2:
            return "SSIMULACRA2"
        if small and config.butteraugli:
            return "Butteraugli"
        for flag, name in ((config.psnr, "PSNR"), (config.dssim, "DSSIM"), (config.ssimulacra2, "SSIMULACRA2"),
                           (config.butteraugli, "Butteraugli")):
            if flag:
                return name
        return "batch"

    def evaluate_batch(self, pairs, config: MetricConfig, intensity_target: float = 80.0) -> List[MetricResult]:
        """Batched calculate_metrics (src/eval/session.rs:437-497).  Raises the first per-pair error,
        like the `?` chain of the reference; use evaluate_batch_raw for per-pair status."""
        out = self.evaluate_batch_raw(pairs, config, intensity_target)
        res = []
        for i, (r, t, w, h) in enumerate(pairs):
            if out[i].status != _lib.CE_OK:
                tsz = np.asarray(t).size
                self._raise(out[i].status, self._failing_metric(config, w, h), (w, h), (tsz // 3 // max(h, 1), h))
            res.append(_result_from_c(out[i]))
        return res

    def sub_batch_capacity(self, config: MetricConfig, width: int, height: int) -> int:
        """Pairs of this size one sub-batch holds in the workspace (ce_sub_batch_capacity)."""
        out = C.c_size_t()
        cfg = config._c()
        self._raise(self._L.ce_sub_batch_capacity(self._h, C.byref(cfg), width, height, C.byref(out)), "batch")
        return int(out.value)

    # -- pinned host memory (opt-in; include/ce_gpu.h "pinned host memory")
    def host_register(self, array: np.ndarray) -> None:
        """Page-lock an existing contiguous array (cudaHostRegister) so its copies run asynchronously at PCIe rate."""
        a = np.asarray(array)
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("host_register needs a C-contiguous array")
        self._raise(self._L.ce_host_register(self._h, C.c_void_p(a.ctypes.data), a.nbytes), "host_register")

    def host_unregister(self, array: np.ndarray) -> None:
        self._raise(self._L.ce_host_unregister(self._h, C.c_void_p(np.asarray(array).ctypes.data)), "host_unregister")

    def evaluate_batch_device(self, d_ref: int, d_dist: int, n: int, width: int, height: int, config: MetricConfig,
                              intensity_target: float = 80.0):
        """Uniform batch already resident in device memory (raw device pointers).  Returns ce_result[n]."""
        out = (_lib.CeResult * max(n, 1))()
        cfg = config._c()
        st = self._L.ce_evaluate_batch_device(self._h, C.c_void_p(d_ref), C.c_void_p(d_dist), n, width, height,
                                              C.byref(cfg), intensity_target, out)
        if st != _lib.CE_OK:
            self._raise(st, "batch")
        return out


    def evaluate_batch_device_grouped(self, d_ref: int, n_ref: int, d_dist: int, n: int, ref_index, width: int, height: int,
                                      config: MetricConfig, intensity_target: float = 80.0):
        """Device-resident batch with shared references: pair i = (reference ref_index[i], distorted i).
        Reference-side work is done once per distinct reference (evaluate_image's shape, session.rs:375-431)."""
        out = (_lib.CeResult * max(n, 1))()
        cfg = config._c()
        ri = np.ascontiguousarray(ref_index, dtype=np.uint32)
        assert ri.size == n
        st = self._L.ce_evaluate_batch_device_grouped(self._h, C.c_void_p(d_ref), n_ref, C.c_void_p(d_dist), n,
                                                      ri.ctypes.data_as(C.POINTER(C.c_uint32)), width, height, C.byref(cfg),
                                                      intensity_target, out)
        if st != _lib.CE_OK:
            self._raise(st, "batch")
        return out


    def transform_to_srgb(self, rgb, width: int, height: int, icc_profile: Optional[bytes]) -> np.ndarray:
        """src/metrics/icc.rs:69-103: RGB8 in the colour space of `icc_profile` -> RGB8 sRGB on the device (matrix/TRC
        profiles).  None / empty = ColorProfile::Srgb (copy).  An unusable profile raises MetricCalculation("ICC", ..)."""
        d = _as_u8(rgb)
        assert d.size == width * height * 3, "Buffer size mismatch"
        out = np.empty(d.size, np.uint8)
        icc = bytes(icc_profile) if icc_profile else b""
        st = self._L.ce_transform_to_srgb(self._h, d.ctypes.data, d.size, width, height, icc if icc else None, len(icc),
                                          out.ctypes.data)
        self._raise(st, "ICC")
        return out

    def transform_profile_to_srgb(self, rgb, profile: ColorProfile) -> np.ndarray:
        """`transform_to_srgb(rgb, &ColorProfile)` (src/metrics/icc.rs:69-103): no dimensions in the signature -- the
        transform is per pixel, so the buffer goes down as one row.  Srgb returns a copy (icc.rs:73)."""
        d = _as_u8(rgb)
        if profile.is_srgb():
            return d.copy()
        if not profile.icc:  # ColorProfile::Icc(vec![]) does not parse (icc.rs:77-81)
            raise MetricCalculation("ICC", "Failed to parse ICC profile: empty profile")
        if d.size % 3:
            raise MetricCalculation("ICC", f"Failed to apply ICC transform: {d.size} bytes is not a whole number of RGB8 pixels")
        if d.size == 0:
            return d.copy()
        out = np.empty(d.size, np.uint8)
        st = self._L.ce_transform_to_srgb(self._h, d.ctypes.data, d.size, d.size // 3, 1, profile.icc, len(profile.icc),
                                          out.ctypes.data)
        self._raise(st, "ICC")
        return out

    def prepare_for_comparison(self, reference, reference_profile: ColorProfile, test, test_profile: ColorProfile):
        """src/metrics/icc.rs:121-130: both images to sRGB, reference first (so its error wins)."""
        return (self.transform_profile_to_srgb(reference, reference_profile),
                self.transform_profile_to_srgb(test, test_profile))

    def calculate_ssimulacra2_icc(self, reference, reference_profile, test, test_profile, width: int, height: int) -> float:
        """src/metrics/ssimulacra2.rs:135-147: ICC transform first, then the plain function (which validates sizes)."""
        r, t = self.prepare_for_comparison(reference, reference_profile, test, test_profile)
        return self.calculate_ssimulacra2(r, t, width, height)

    def calculate_butteraugli_icc(self, reference, reference_profile, test, test_profile, width: int, height: int) -> float:
        """src/metrics/butteraugli.rs:150-162."""
        r, t = self.prepare_for_comparison(reference, reference_profile, test, test_profile)
        return self.calculate_butteraugli(r, t, width, height)

    def calculate_dssim_icc(self, reference, reference_profile, test, test_profile, width: int, height: int,
                            viewing=None) -> float:
        """src/metrics/dssim.rs:158-174: transform, rgb8_to_dssim_image x2, calculate_dssim (fused on the device)."""
        r, t = self.prepare_for_comparison(reference, reference_profile, test, test_profile)
        return self.calculate_dssim_rgb8(r, t, width, height)

    # -- on-device distortion source (the step before the metric path in codec-iter's run_eval,
    #    crates/codec-iter/src/eval.rs:153-172)
    def jpeg_roundtrip(self, rgb, width: int, height: int, quality: int, subsampling: int = 2) -> np.ndarray:
        """Decoded image of a baseline JPEG(quality, subsampling 0 = 4:4:4 / 2 = 4:2:0) of `rgb`, computed on the
        device; bit-exact with libjpeg-turbo's encode -> decode.  uint8 [height, width, 3]."""
        d = _as_u8(rgb)
        assert d.size == width * height * 3, "Buffer size mismatch"
        out = np.empty((height, width, 3), np.uint8)
        st = self._L.ce_jpeg_roundtrip(self._h, d.ctypes.data, d.size, width, height, int(quality), int(subsampling),
                                       out.ctypes.data)
        self._raise(st, "JPEG")
        return out

    def jpeg_roundtrip_device(self, d_refs: int, n_ref: int, width: int, height: int, qualities, subsampling: int, d_out: int):
        """Device-resident: d_out[r * len(qualities) + k] = reference r at qualities[k] (raw device pointers)."""
        q = (C.c_int * len(qualities))(*[int(v) for v in qualities])
        st = self._L.ce_jpeg_roundtrip_device(self._h, C.c_void_p(d_refs), n_ref, width, height, q, len(qualities),
                                              int(subsampling), C.c_void_p(d_out))
        self._raise(st, "JPEG")

    def evaluate_jpeg_sweep_raw(self, refs, width: int, height: int, qualities, config: MetricConfig, subsampling: int = 2,
                                intensity_target: float = 80.0):
        """Quality sweep: every reference against its own JPEG round trip at every quality, distortions generated on
        the device (only the references are uploaded).  Returns ce_result[len(refs) * len(qualities)], reference-major."""
        keep = [_as_u8(r) for r in refs]
        for r in keep:
            assert r.size == width * height * 3, "Buffer size mismatch"
        n_ref, n_q = len(keep), len(qualities)
        ptrs = (C.c_void_p * max(n_ref, 1))(*[r.ctypes.data for r in keep])
        q = (C.c_int * max(n_q, 1))(*[int(v) for v in qualities])
        out = (_lib.CeResult * max(n_ref * n_q, 1))()
        cfg = config._c()
        st = self._L.ce_evaluate_jpeg_sweep(self._h, ptrs, n_ref, width, height, q, n_q, int(subsampling), C.byref(cfg),
                                            intensity_target, out)
        if st != _lib.CE_OK:
            self._raise(st, "batch")
        return out

    def evaluate_jpeg_sweep(self, refs, width: int, height: int, qualities, config: MetricConfig, subsampling: int = 2,
                            intensity_target: float = 80.0) -> List[List[MetricResult]]:
        """[reference][quality] MetricResult table of evaluate_jpeg_sweep_raw."""
        out = self.evaluate_jpeg_sweep_raw(refs, width, height, qualities, config, subsampling, intensity_target)
        n_q = len(qualities)
        table = []
        for r in range(len(refs)):
            row = []
            for k in range(n_q):
                o = out[r * n_q + k]
                if o.status != _lib.CE_OK:
                    self._raise(o.status, self._failing_metric(config, width, height), (width, height), (width, height))
                row.append(_result_from_c(o))
            table.append(row)
        return table


class GpuReference:
    """Reference image kept on the device; mirrors fast_ssim2::Ssimulacra2Reference::new / .compare
    as used by crates/codec-iter/src/eval.rs:138-149,84-88."""

    def __init__(self, ctx: GpuMetrics, reference, width: int, height: int, config: MetricConfig):
        self._ctx = ctx
        self._config = config
        r = _as_u8(reference)
        h = C.c_void_p()
        cfg = config._c()
        st = ctx._L.ce_reference_create(ctx._h, r.ctypes.data, r.size, width, height, C.byref(cfg), C.byref(h))
        ctx._raise(st, "SSIMULACRA2")
        self._h, self.width, self.height = h, width, height
        ctx._refs.add(self)

    def compare(self, test, intensity_target: float = 80.0) -> MetricResult:
        return self.compare_many([test], intensity_target)[0]

    def compare_many(self, tests, intensity_target: float = 80.0) -> List[MetricResult]:
        ts = [_as_u8(t) for t in tests]
        n = len(ts)
        ptrs = (C.c_void_p * n)(*[t.ctypes.data for t in ts])
        lens = (C.c_size_t * n)(*[t.size for t in ts])
        out = (_lib.CeResult * n)()
        st = self._ctx._L.ce_reference_compare_many(self._ctx._h, self._h, ptrs, lens, n, intensity_target, out)
        self._ctx._raise(st, "batch")
        res = []
        for i in range(n):
            self._ctx._raise(out[i].status, self._ctx._failing_metric(self._config, self.width, self.height), (self.width, self.height),
                             (ts[i].size // 3 // max(self.height, 1), self.height))
            res.append(_result_from_c(out[i]))
        return res

    def close(self):
        if getattr(self, "_h", None):
            self._ctx._L.ce_reference_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- module-level functions
_default: Optional[GpuMetrics] = None


def default_context() -> GpuMetrics:
    global _default
    if _default is None:
        _default = GpuMetrics(0)
    return _default


def calculate_psnr(reference, test, width, height) -> float:
    return default_context().calculate_psnr(reference, test, width, height)


def calculate_ssimulacra2(reference, test, width, height) -> float:
    return default_context().calculate_ssimulacra2(reference, test, width, height)


def calculate_butteraugli(reference, test, width, height) -> float:
    return default_context().calculate_butteraugli(reference, test, width, height)


def calculate_butteraugli_with_intensity(reference, test, width, height, intensity_target) -> float:
    return default_context().calculate_butteraugli_with_intensity(reference, test, width, height, intensity_target)


def calculate_dssim(reference, test, viewing=None) -> float:
    return default_context().calculate_dssim(reference, test, viewing)


def rgb8_to_dssim_image(data, width, height) -> np.ndarray:
    return default_context().rgb8_to_dssim_image(data, width, height)


def rgba8_to_dssim_image(data, width, height) -> np.ndarray:
    return default_context().rgba8_to_dssim_image(data, width, height)


def xyb_roundtrip(rgb, width, height) -> np.ndarray:
    return default_context().xyb_roundtrip(rgb, width, height)


def transform_to_srgb(rgb, profile: ColorProfile) -> np.ndarray:
    return default_context().transform_profile_to_srgb(rgb, profile)


def prepare_for_comparison(reference, reference_profile, test, test_profile):
    return default_context().prepare_for_comparison(reference, reference_profile, test, test_profile)


def calculate_ssimulacra2_icc(reference, reference_profile, test, test_profile, width, height) -> float:
    return default_context().calculate_ssimulacra2_icc(reference, reference_profile, test, test_profile, width, height)


def calculate_butteraugli_icc(reference, reference_profile, test, test_profile, width, height) -> float:
    return default_context().calculate_butteraugli_icc(reference, reference_profile, test, test_profile, width, height)


def calculate_dssim_icc(reference, reference_profile, test, test_profile, width, height, viewing=None) -> float:
    return default_context().calculate_dssim_icc(reference, reference_profile, test, test_profile, width, height, viewing)
