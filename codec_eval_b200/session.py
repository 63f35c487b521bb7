"""Host-side mirror of codec-eval's `src/eval` caller of the metric path:

  ImageData / EncodeRequest / EvalConfig(+builder) / EvalSession        src/eval/session.rs:25-66,160-278,302-523
  CodecResult / ImageReport / CorpusReport                              src/eval/report.rs:16-182
  write_image_report / write_corpus_report (JSON + 13-column CSV)       src/eval/session.rs:500-584

The one behavioural change is the dispatch the north star asks for: `evaluate_image` no longer calls
`calculate_metrics` once per (codec, quality) (session.rs:375-431); it encodes/decodes everything first, then issues
ONE batched GPU call for all decoded outputs of the reference (shared host pointer => the reference is uploaded and
pre-processed once), and fills `CodecResult.metrics` / `.perception` in the original order.  Metric arithmetic runs
only in the CUDA library (codec_eval_b200.metrics.GpuMetrics); this file marshals buffers and formats reports.
"""
from __future__ import annotations

import csv
import datetime as _dt
import json
import math
import os
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .metrics import MetricConfig, MetricResult, PerceptionLevel


# ----------------------------------------------------------------------------- ImageData (session.rs:25-147)
class ImageData:
    """RGB8 / RGBA8 pixels, row-major.  Variants of the reference enum: Rgb8, Rgba8, RgbSlice, RgbaSlice,
    RgbSliceWithIcc."""

    def __init__(self, data, width: int, height: int, channels: int, icc_profile: Optional[bytes] = None):
        a = np.ascontiguousarray(np.asarray(data, dtype=np.uint8)).reshape(-1)
        if a.size != width * height * channels:
            raise ValueError(f"expected {width * height * channels} bytes, got {a.size}")
        self.data, self._w, self._h, self.channels, self.icc_profile = a, int(width), int(height), channels, icc_profile

    @staticmethod
    def rgb8(img: np.ndarray) -> "ImageData":                       # ImageData::Rgb8(ImgVec<RGB8>)
        return ImageData(img, img.shape[1], img.shape[0], 3)

    @staticmethod
    def rgba8(img: np.ndarray) -> "ImageData":                      # ImageData::Rgba8(ImgVec<RGBA8>)
        return ImageData(img, img.shape[1], img.shape[0], 4)

    @staticmethod
    def rgb_slice(data, width: int, height: int) -> "ImageData":
        return ImageData(data, width, height, 3)

    @staticmethod
    def rgba_slice(data, width: int, height: int) -> "ImageData":
        return ImageData(data, width, height, 4)

    @staticmethod
    def rgb_slice_with_icc(data, width: int, height: int, icc_profile: bytes) -> "ImageData":
        return ImageData(data, width, height, 3, icc_profile)

    def width(self) -> int:
        return self._w

    def height(self) -> int:
        return self._h

    def to_rgb8_vec(self) -> np.ndarray:
        """Tight RGB8 (alpha dropped), no ICC transform (session.rs:98-117)."""
        if self.channels == 3:
            return self.data
        return np.ascontiguousarray(self.data.reshape(-1, 4)[:, :3]).reshape(-1)

    def to_rgb8_srgb(self, metrics=None) -> np.ndarray:
        """session.rs:143-147: to_rgb8_vec, then the ICC -> sRGB transform (src/metrics/icc.rs:69-103) when a profile is
        attached -- here on the device (matrix/TRC profiles; others raise MetricCalculation("ICC", ..))."""
        rgb = self.to_rgb8_vec()
        if not self.icc_profile:
            return rgb
        from .metrics import default_context

        ctx = metrics or default_context()
        return ctx.transform_to_srgb(rgb, self.width(), self.height(), self.icc_profile)


@dataclass
class EncodeRequest:  # session.rs:160-176
    quality: float
    params: Dict[str, str] = field(default_factory=dict)

    def with_param(self, key: str, value: str) -> "EncodeRequest":
        self.params[key] = value
        return self


# ----------------------------------------------------------------------------- reports (report.rs)
@dataclass
class CodecResult:
    codec_id: str
    codec_version: str
    quality: float
    file_size: int
    bits_per_pixel: float
    encode_time: float                      # seconds (Duration); serialised as integer milliseconds
    decode_time: Optional[float]
    metrics: MetricResult
    perception: Optional[PerceptionLevel]
    cached_path: Optional[str] = None
    codec_params: Dict[str, str] = field(default_factory=dict)

    def compression_ratio(self, original_size: int) -> float:
        return 0.0 if self.file_size == 0 else original_size / self.file_size

    def to_json(self) -> dict:
        m = self.metrics
        return {
            "codec_id": self.codec_id, "codec_version": self.codec_version, "quality": float(self.quality),
            "file_size": self.file_size, "bits_per_pixel": self.bits_per_pixel,
            "encode_time": int(self.encode_time * 1000), "decode_time": None if self.decode_time is None else int(self.decode_time * 1000),
            # serde_json writes a non-finite f64 as null (PSNR of identical images is +inf, src/metrics/mod.rs:326-328);
            # Python's bare `Infinity` token is not JSON and the reference's reader rejects it
            "metrics": {"dssim": _json_f64(m.dssim), "ssimulacra2": _json_f64(m.ssimulacra2),
                        "butteraugli": _json_f64(m.butteraugli), "psnr": _json_f64(m.psnr)},
            "perception": None if self.perception is None else self.perception.name,
            "cached_path": self.cached_path, "codec_params": dict(self.codec_params),
        }


def _json_f64(v: Optional[float]) -> Optional[float]:
    return None if v is None or not math.isfinite(v) else v


def _now() -> str:
    return _dt.datetime.now(_dt.timezone.utc).isoformat()   # RFC 3339, like chrono's to_rfc3339


@dataclass
class ImageReport:
    name: str
    width: int
    height: int
    source_path: Optional[str] = None
    uncompressed_size: int = 0
    results: List[CodecResult] = field(default_factory=list)
    timestamp: str = field(default_factory=_now)

    def __post_init__(self):
        if not self.uncompressed_size:
            self.uncompressed_size = self.width * self.height * 3

    def results_for_codec(self, codec_id: str):
        return (r for r in self.results if r.codec_id == codec_id)

    def best_at_size(self, max_bytes: int) -> Optional[CodecResult]:
        c = [r for r in self.results if r.file_size <= max_bytes]
        if not c:
            return None
        return max(c, key=lambda r: -r.metrics.dssim if r.metrics.dssim is not None else float("-inf"))

    def smallest_at_quality(self, max_dssim: float) -> Optional[CodecResult]:
        c = [r for r in self.results if r.metrics.dssim is not None and r.metrics.dssim <= max_dssim]
        return min(c, key=lambda r: r.file_size) if c else None

    def to_json(self) -> dict:
        return {"name": self.name, "source_path": self.source_path, "width": self.width, "height": self.height,
                "uncompressed_size": self.uncompressed_size, "results": [r.to_json() for r in self.results],
                "timestamp": self.timestamp}


@dataclass
class CorpusReport:
    name: str
    images: List[ImageReport] = field(default_factory=list)
    timestamp: str = field(default_factory=_now)
    config_summary: str = ""

    def total_results(self) -> int:
        return sum(len(i.results) for i in self.images)

    def codec_ids(self) -> List[str]:
        return sorted({r.codec_id for i in self.images for r in i.results})

    def to_json(self) -> dict:
        return {"name": self.name, "images": [i.to_json() for i in self.images], "timestamp": self.timestamp,
                "config_summary": self.config_summary}


# ----------------------------------------------------------------------------- config (session.rs:189-278)
@dataclass
class EvalConfig:
    report_dir: str
    cache_dir: Optional[str] = None
    viewing: object = None
    metrics: MetricConfig = field(default_factory=MetricConfig.all)
    quality_levels: List[float] = field(default_factory=lambda: [50.0, 60.0, 70.0, 80.0, 85.0, 90.0, 95.0])

    @staticmethod
    def builder() -> "EvalConfigBuilder":
        return EvalConfigBuilder()


class EvalConfigBuilder:
    def __init__(self):
        self._kw = {}

    def report_dir(self, p):
        self._kw["report_dir"] = str(p)
        return self

    def cache_dir(self, p):
        self._kw["cache_dir"] = str(p)
        return self

    def viewing(self, v):
        self._kw["viewing"] = v
        return self

    def metrics(self, m: MetricConfig):
        self._kw["metrics"] = m
        return self

    def quality_levels(self, levels: Sequence[float]):
        self._kw["quality_levels"] = [float(x) for x in levels]
        return self

    def build(self) -> EvalConfig:
        if "report_dir" not in self._kw:
            raise AssertionError("report_dir is required")   # .expect("report_dir is required"), session.rs:268
        return EvalConfig(**self._kw)


EncodeFn = Callable[[ImageData, EncodeRequest], bytes]
DecodeFn = Callable[[bytes], ImageData]


@dataclass
class _CodecEntry:
    id: str
    version: str
    encode: EncodeFn
    decode: Optional[DecodeFn]


def _fmt_f64(x: float) -> str:
    """Rust's `f64::to_string`: shortest round-trip repr without a trailing `.0`."""
    return repr(int(x)) if float(x).is_integer() else repr(float(x))


# ----------------------------------------------------------------------------- the session
class EvalSession:
    """src/eval/session.rs:302-523.  `metrics` is the batched metric backend: anything with
    `evaluate_batch(pairs, MetricConfig) -> list[MetricResult]` (default: a GpuMetrics context on device 0)."""

    def __init__(self, config: EvalConfig, metrics=None):
        self.config = config
        self.codecs: List[_CodecEntry] = []
        self._metrics = metrics

    def _backend(self):
        if self._metrics is None:
            from .metrics import GpuMetrics

            self._metrics = GpuMetrics(0)
        return self._metrics

    def add_codec(self, id: str, version: str, encode: EncodeFn) -> "EvalSession":
        self.codecs.append(_CodecEntry(id, version, encode, None))
        return self

    def add_codec_with_decode(self, id: str, version: str, encode: EncodeFn, decode: DecodeFn) -> "EvalSession":
        self.codecs.append(_CodecEntry(id, version, encode, decode))
        return self

    def codec_count(self) -> int:
        return len(self.codecs)

    def evaluate_image(self, name: str, image: ImageData) -> ImageReport:
        width, height = image.width(), image.height()
        report = ImageReport(name, width, height)
        reference_rgb = image.to_rgb8_vec()
        pending = []   # (index into report.results, decoded rgb)
        for codec in self.codecs:
            for quality in self.config.quality_levels:
                request = EncodeRequest(quality)
                t0 = time.perf_counter()
                encoded = codec.encode(image, request)
                encode_time = time.perf_counter() - t0
                decode_time = None
                if codec.decode is not None:
                    t0 = time.perf_counter()
                    decoded = codec.decode(encoded)
                    decode_time = time.perf_counter() - t0
                    decoded_rgb = decoded.to_rgb8_srgb()
                    pending.append((len(report.results), decoded_rgb, decoded.width(), decoded.height()))
                report.results.append(CodecResult(
                    codec_id=codec.id, codec_version=codec.version, quality=quality, file_size=len(encoded),
                    bits_per_pixel=(len(encoded) * 8) / (float(width) * float(height)), encode_time=encode_time,
                    decode_time=decode_time, metrics=MetricResult(), perception=None, cached_path=None,
                    codec_params=dict(request.params)))
        if pending:
            # one batched call; every pair shares the reference buffer, so it is uploaded and prepared once.
            # The buffer lengths drive the same validation as calculate_metrics (a decoder that returns another
            # size surfaces as DimensionMismatch / MetricCalculation, first failing pair first).
            pairs = [(reference_rgb, d, width, height) for _, d, _, _ in pending]
            results = self._backend().evaluate_batch(pairs, self.config.metrics)
            for (idx, _, _, _), m in zip(pending, results):
                m = MetricResult(dssim=m.dssim, ssimulacra2=m.ssimulacra2, butteraugli=m.butteraugli, psnr=m.psnr)
                report.results[idx].metrics = m
                report.results[idx].perception = m.perception_level()
        return report

    def evaluate_corpus(self, name: str, images: Sequence) -> CorpusReport:
        """Convenience: images = [(name, ImageData)]."""
        rep = CorpusReport(name)
        rep.config_summary = f"metrics={self.config.metrics} quality_levels={self.config.quality_levels}"
        for n, img in images:
            rep.images.append(self.evaluate_image(n, img))
        return rep

    # -- writers (session.rs:500-584)
    def write_image_report(self, report: ImageReport) -> None:
        os.makedirs(self.config.report_dir, exist_ok=True)
        with open(os.path.join(self.config.report_dir, f"{report.name}.json"), "w") as f:
            json.dump(report.to_json(), f, indent=2, allow_nan=False)

    def write_corpus_report(self, report: CorpusReport) -> None:
        os.makedirs(self.config.report_dir, exist_ok=True)
        with open(os.path.join(self.config.report_dir, f"{report.name}.json"), "w") as f:
            json.dump(report.to_json(), f, indent=2, allow_nan=False)
        self.write_csv_summary(report, os.path.join(self.config.report_dir, f"{report.name}.csv"))

    @staticmethod
    def write_csv_summary(report: CorpusReport, path: str) -> None:
        with open(path, "w", newline="") as f:
            w = csv.writer(f, lineterminator="\n")
            w.writerow(["image", "codec", "version", "quality", "file_size", "bpp", "encode_ms", "decode_ms", "dssim",
                        "ssimulacra2", "butteraugli", "psnr", "perception"])
            for img in report.images:
                for r in img.results:
                    m = r.metrics
                    w.writerow([
                        img.name, r.codec_id, r.codec_version, _fmt_f64(r.quality), str(r.file_size), f"{r.bits_per_pixel:.4f}",
                        str(int(r.encode_time * 1000)), "" if r.decode_time is None else str(int(r.decode_time * 1000)),
                        "" if m.dssim is None else f"{m.dssim:.6f}", "" if m.ssimulacra2 is None else f"{m.ssimulacra2:.2f}",
                        "" if m.butteraugli is None else f"{m.butteraugli:.4f}", "" if m.psnr is None else f"{m.psnr:.2f}",
                        "" if r.perception is None else r.perception.code(),
                    ])


# ----------------------------------------------------------------------------- helpers (src/eval/helpers.rs)
def evaluate_single(reference: np.ndarray, encoded: np.ndarray, config: MetricConfig, metrics=None) -> MetricResult:
    """helpers.rs:105-173: [h, w, 3] uint8 images (ImgVec<RGB8>; may be strided views)."""
    from .metrics import DimensionMismatch, default_context

    if reference.shape[:2] != encoded.shape[:2]:
        raise DimensionMismatch((reference.shape[1], reference.shape[0]), (encoded.shape[1], encoded.shape[0]))
    h, w = reference.shape[:2]
    ctx = metrics or default_context()
    r = ctx.evaluate_batch([(np.ascontiguousarray(reference), np.ascontiguousarray(encoded), w, h)], config)[0]
    return MetricResult(dssim=r.dssim, ssimulacra2=r.ssimulacra2, butteraugli=r.butteraugli, psnr=r.psnr)


class QualityBelowThreshold(Exception):  # src/error.rs:68
    def __init__(self, metric: str, value: float, threshold: float):
        super().__init__(f"Quality below threshold: {metric} = {value}, threshold = {threshold}")
        self.metric, self.value, self.threshold = metric, value, threshold


def assert_quality(reference: np.ndarray, encoded: np.ndarray, min_ssimulacra2: Optional[float], max_dssim: Optional[float],
                   metrics=None) -> MetricResult:
    """helpers.rs:212-255."""
    cfg = MetricConfig(dssim=max_dssim is not None, ssimulacra2=min_ssimulacra2 is not None)
    r = evaluate_single(reference, encoded, cfg, metrics)
    if min_ssimulacra2 is not None and r.ssimulacra2 is not None and r.ssimulacra2 < min_ssimulacra2:
        raise QualityBelowThreshold("SSIMULACRA2", r.ssimulacra2, min_ssimulacra2)
    if max_dssim is not None and r.dssim is not None and r.dssim > max_dssim:
        raise QualityBelowThreshold("DSSIM", r.dssim, max_dssim)
    return r


_LEVEL_ORDER = [PerceptionLevel.Imperceptible, PerceptionLevel.Marginal, PerceptionLevel.Subtle, PerceptionLevel.Noticeable,
                PerceptionLevel.Degraded]


def assert_perception_level(reference: np.ndarray, encoded: np.ndarray, min_level: PerceptionLevel, metrics=None) -> MetricResult:
    """helpers.rs:291-321: the DSSIM perception level must be `min_level` or better (enum ordinals compared)."""
    r = evaluate_single(reference, encoded, MetricConfig(dssim=True), metrics)
    if r.dssim is not None:
        actual = _LEVEL_ORDER.index(PerceptionLevel.from_dssim(r.dssim))
        wanted = _LEVEL_ORDER.index(min_level)
        if actual > wanted:
            raise QualityBelowThreshold(f"PerceptionLevel (DSSIM {r.dssim:.6f})", float(actual), float(wanted))
    return r
