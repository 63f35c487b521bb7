"""Host-side twin of codec-iter's eval loop and its on-disk formats, so the reference's own
tools consume GPU-produced scores unchanged:

* source cache : `<corpus>/.codec-iter-cache/<stem>.ppm` (binary P6), crates/codec-iter/src/source.rs:62-92,176-201
* EvalPoint / EvalResult : crates/codec-iter/src/eval.rs:21-35
* Baseline JSON (`<baselines>/<format>.json`), compare_with_baseline, aggregate_by_quality :
  crates/codec-iter/src/baseline.rs:12-104
* run_eval : crates/codec-iter/src/eval.rs:93-215 -- here every (image, quality) output of a codec is decoded first and
  all of one image's distortions go to the GPU in ONE grouped call (the reference compares them one at a time against a
  precomputed Ssimulacra2Reference, eval.rs:138-149,190-196); `run_eval_jpeg_sweep` keeps even the decoded images on
  the device (on-device baseline-JPEG source, include/ce_gpu.h ce_evaluate_jpeg_sweep).

The metric itself always runs on the GPU through codec_eval_b200.metrics -- there is no CPU fallback here either.
"""
from __future__ import annotations

import datetime
import json
import os
import time
from dataclasses import asdict, dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .metrics import GpuMetrics, MetricConfig, default_context

CACHE_DIR = ".codec-iter-cache"


# ------------------------------------------------------------------ sources (source.rs)
@dataclass
class SourceImage:  # source.rs:10-15
    name: str
    width: int
    height: int
    pixels: np.ndarray  # uint8 [height, width, 3]


def encode_ppm(pixels: np.ndarray) -> bytes:
    """Binary PPM (P6, maxval 255), what zenbitmaps::encode_ppm_img writes (source.rs:193-201)."""
    p = np.ascontiguousarray(pixels, dtype=np.uint8)
    assert p.ndim == 3 and p.shape[2] == 3
    return b"P6\n%d %d\n255\n" % (p.shape[1], p.shape[0]) + p.tobytes()


def decode_ppm(data: bytes) -> np.ndarray:
    """Binary P6 with maxval <= 255; header tokens separated by any whitespace, `#` comments allowed."""
    pos, tokens = 0, []
    while len(tokens) < 4:
        while pos < len(data) and data[pos:pos + 1].isspace():
            pos += 1
        if pos >= len(data):
            raise ValueError("truncated PPM header")
        if data[pos:pos + 1] == b"#":
            while pos < len(data) and data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while pos < len(data) and not data[pos:pos + 1].isspace():
            pos += 1
        tokens.append(data[start:pos])
    if tokens[0] != b"P6":
        raise ValueError(f"not a binary PPM (magic {tokens[0]!r})")
    w, h, maxval = int(tokens[1]), int(tokens[2]), int(tokens[3])
    if w <= 0 or h <= 0:
        raise ValueError(f"bad PPM dimensions {w}x{h}")
    if not 0 < maxval <= 255:
        raise ValueError("only 8-bit PPM is supported")
    pos += 1  # the single whitespace byte after maxval
    need = w * h * 3
    if len(data) - pos < need:
        raise ValueError("truncated PPM data")
    return np.frombuffer(data, np.uint8, need, pos).reshape(h, w, 3).copy()


def load_ppm(path: str, name: str) -> SourceImage:  # source.rs:176-191
    px = decode_ppm(open(path, "rb").read())
    return SourceImage(name, px.shape[1], px.shape[0], px)


def cache_as_ppm(img: SourceImage, cache_dir: str, name: str) -> None:  # source.rs:193-203
    os.makedirs(cache_dir, exist_ok=True)
    stem = os.path.splitext(os.path.basename(name))[0]
    with open(os.path.join(cache_dir, stem + ".ppm"), "wb") as f:
        f.write(encode_ppm(img.pixels))


def load_png(path: str, name: str) -> SourceImage:  # source.rs:126-174 (8-bit RGB; alpha dropped, grey expanded)
    from PIL import Image

    px = np.asarray(Image.open(path).convert("RGB")).copy()
    return SourceImage(name, px.shape[1], px.shape[0], px)


def load_by_names(corpus: str, names: Sequence[str]) -> List[SourceImage]:  # source.rs:62-92
    cache_dir = os.path.join(corpus, CACHE_DIR)
    images = []
    for name in names:
        stem = os.path.splitext(os.path.basename(name))[0]
        ppm_path, png_path = os.path.join(cache_dir, stem + ".ppm"), os.path.join(corpus, name)
        if os.path.exists(ppm_path):
            img = load_ppm(ppm_path, name)
        elif os.path.exists(png_path):
            img = load_png(png_path, name)
            try:
                cache_as_ppm(img, cache_dir, name)
            except OSError as e:  # the reference only warns
                print(f"warning: failed to cache PPM for {name}: {e}")
        else:
            raise FileNotFoundError(f"Image not found: {name} (looked in {png_path} and {ppm_path})")
        images.append(img)
    return images


def load_all_from_dir(corpus: str, limit: int = 0) -> List[SourceImage]:  # source.rs:94-124: sorted *.png, first `limit`
    names = sorted(n for n in os.listdir(corpus) if n.lower().endswith(".png"))
    if limit:
        names = names[:limit]
    return load_by_names(corpus, names)


# ------------------------------------------------------------------ eval points / baseline (eval.rs, baseline.rs)
@dataclass
class EvalPoint:  # eval.rs:21-29, field order = serde order
    image: str
    quality: int
    bpp: float
    ssim2: float
    size_bytes: int
    encode_ms: int


@dataclass
class EvalResult:  # eval.rs:31-35
    config_summary: str
    points: List[EvalPoint]
    total_ms: int


def _now_rfc3339() -> str:
    return datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.%fZ")


@dataclass
class Baseline:  # baseline.rs:12-18
    format: str
    config_summary: str
    corpus_path: str
    created_at: str = field(default_factory=_now_rfc3339)
    points: List[EvalPoint] = field(default_factory=list)

    def to_json(self) -> str:  # serde_json::to_string_pretty: 2-space indent, struct field order
        return json.dumps({"format": self.format, "config_summary": self.config_summary, "corpus_path": self.corpus_path,
                           "created_at": self.created_at, "points": [asdict(p) for p in self.points]}, indent=2)

    @staticmethod
    def from_json(text: str) -> "Baseline":
        d = json.loads(text)
        return Baseline(d["format"], d["config_summary"], d["corpus_path"], d["created_at"],
                        [EvalPoint(p["image"], int(p["quality"]), float(p["bpp"]), float(p["ssim2"]), int(p["size_bytes"]),
                                   int(p["encode_ms"])) for p in d["points"]])


def baseline_path(baselines_dir: str, fmt: str) -> str:  # baseline.rs:20-22
    return os.path.join(baselines_dir, f"{fmt}.json")


def load_baseline(baselines_dir: str, fmt: str) -> Optional[Baseline]:  # baseline.rs:24-34
    p = baseline_path(baselines_dir, fmt)
    if not os.path.exists(p):
        return None
    return Baseline.from_json(open(p).read())


def save_baseline(baselines_dir: str, baseline: Baseline) -> None:  # baseline.rs:36-43
    os.makedirs(baselines_dir, exist_ok=True)
    with open(baseline_path(baselines_dir, baseline.format), "w") as f:
        f.write(baseline.to_json())


@dataclass
class ComparisonRow:  # baseline.rs:45-52
    quality: int
    bpp: float
    ssim2: float
    delta_bpp: float
    delta_ssim2: float
    pareto: float


def aggregate_by_quality(points: Sequence[EvalPoint]) -> Dict[int, Tuple[float, float]]:  # baseline.rs:89-104
    acc: Dict[int, Tuple[List[float], List[float]]] = {}
    for p in points:
        b, s = acc.setdefault(p.quality, ([], []))
        b.append(p.bpp)
        s.append(p.ssim2)
    out = {}
    for q, (b, s) in acc.items():
        n = float(len(b))
        sb = ss = 0.0
        for v in b:  # f64 sums in submission order, as iter().sum()
            sb += v
        for v in s:
            ss += v
        out[q] = (sb / n, ss / n)
    return out


def compare_with_baseline(points: Sequence[EvalPoint], baseline: Baseline) -> List[ComparisonRow]:  # baseline.rs:54-87
    cur, base = aggregate_by_quality(points), aggregate_by_quality(baseline.points)
    rows = []
    for q in sorted(cur):
        bpp, s2 = cur[q]
        d_bpp, d_s2 = (bpp - base[q][0], s2 - base[q][1]) if q in base else (0.0, 0.0)
        rows.append(ComparisonRow(q, bpp, s2, d_bpp, d_s2, d_s2 - d_bpp * 10.0))
    return rows


# ------------------------------------------------------------------ run_eval (eval.rs:93-215)
@dataclass
class CodecConfig:  # eval.rs:10-19
    encode: Callable[[np.ndarray, int], bytes]     # (pixels [h,w,3] u8, quality) -> compressed bytes
    decode: Callable[[bytes], np.ndarray]          # compressed bytes -> pixels [h,w,3] u8
    summary: str


def run_eval(images: Sequence[SourceImage], quality_levels: Sequence[int], codec: CodecConfig,
             metrics: Optional[GpuMetrics] = None) -> EvalResult:
    """codec-iter's run_eval with the comparisons of one image batched: encode + decode every quality on the host (the
    codec is the caller's), then ONE grouped GPU call for all of that image's distortions (SSIMULACRA2 only, as the
    reference).  Point order is the reference's: image-major, quality-minor."""
    ctx = metrics or default_context()
    t0 = time.perf_counter()
    cfg = MetricConfig.ssimulacra2_only()
    points: List[EvalPoint] = []
    for image in images:
        total_pixels = float(image.width * image.height)
        meta, pairs = [], []
        for q in quality_levels:
            e0 = time.perf_counter()
            try:
                encoded = codec.encode(image.pixels, int(q))
            except Exception as e:  # eval.rs:175-176
                raise RuntimeError(f"Encode error for {image.name} q{q}: {e}") from e
            encode_ms = int((time.perf_counter() - e0) * 1000)
            try:
                decoded = np.ascontiguousarray(codec.decode(encoded), dtype=np.uint8)
            except Exception as e:  # eval.rs:183-184
                raise RuntimeError(f"Decode error for {image.name} q{q}: {e}") from e
            meta.append((int(q), len(encoded), encode_ms))
            pairs.append((image.pixels, decoded, image.width, image.height))
        # all pairs share the reference buffer => uploaded and preprocessed once (Ssimulacra2Reference reuse)
        res = ctx.evaluate_batch(pairs, cfg)
        for (q, size_bytes, encode_ms), r in zip(meta, res):
            points.append(EvalPoint(image.name, q, size_bytes * 8.0 / total_pixels, r.ssimulacra2, size_bytes, encode_ms))
    return EvalResult(codec.summary, points, int((time.perf_counter() - t0) * 1000))


def run_eval_jpeg_sweep(images: Sequence[SourceImage], quality_levels: Sequence[int], subsampling: int = 2,
                        size_of: Optional[Callable[[np.ndarray, int], int]] = None,
                        metrics: Optional[GpuMetrics] = None) -> EvalResult:
    """The same sweep for baseline JPEG with the decoded images generated ON the device: only the sources are uploaded
    (one call per image size).  The device source has no bitstream, so `size_of(pixels, quality)` supplies the file
    size (e.g. a real encoder run without decoding); without it bpp / size_bytes are 0."""
    ctx = metrics or default_context()
    t0 = time.perf_counter()
    cfg = MetricConfig.ssimulacra2_only()
    by_size: Dict[Tuple[int, int], List[int]] = {}
    for i, im in enumerate(images):
        by_size.setdefault((im.width, im.height), []).append(i)
    scores: Dict[int, List[float]] = {}
    for (w, h), idx in by_size.items():
        table = ctx.evaluate_jpeg_sweep([images[i].pixels for i in idx], w, h, list(quality_levels), cfg, subsampling)
        for i, row in zip(idx, table):
            scores[i] = [r.ssimulacra2 for r in row]
    points: List[EvalPoint] = []
    for i, im in enumerate(images):
        for k, q in enumerate(quality_levels):
            e0 = time.perf_counter()
            size_bytes = int(size_of(im.pixels, int(q))) if size_of else 0
            encode_ms = int((time.perf_counter() - e0) * 1000) if size_of else 0
            points.append(EvalPoint(im.name, int(q), size_bytes * 8.0 / float(im.width * im.height), scores[i][k], size_bytes,
                                    encode_ms))
    ss = "444" if subsampling == 0 else "420"
    return EvalResult(f"ce-gpu-jpeg-{ss}-ycbcr-baseline", points, int((time.perf_counter() - t0) * 1000))
