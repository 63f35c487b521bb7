"""The caller of the metric path: EvalSession batching, result order, report formats.
CPU tests use a recording fake metric backend (host logic only); the GPU test runs the real CUDA path."""
import csv
import json
import os

import numpy as np
import pytest

from codec_eval_b200.metrics import MetricConfig, MetricResult, PerceptionLevel
from codec_eval_b200.session import (CodecResult, CorpusReport, EncodeRequest, EvalConfig, EvalSession, ImageData, ImageReport,
                                     QualityBelowThreshold)


class FakeMetrics:
    """Records the batched calls; the "score" encodes which decoded buffer it saw."""

    def __init__(self):
        self.calls = []

    def evaluate_batch(self, pairs, config):
        self.calls.append((len(pairs), {id(p[0]) for p in pairs}, config))
        out = []
        for ref, dist, w, h in pairs:
            q = float(dist[0])      # the fake decoder writes the quality into byte 0
            out.append(MetricResult(dssim=(100 - q) * 1e-5, ssimulacra2=q, butteraugli=(100 - q) / 20, psnr=20 + q / 5))
        return out


def _img(w=16, h=8):
    return ImageData.rgb_slice((np.arange(w * h * 3) % 256).astype(np.uint8), w, h)


def _codec(tag):
    def enc(image, req):
        return bytes([int(req.quality)]) * (10 + int(req.quality) + tag)

    def dec(data):
        px = np.zeros(16 * 8 * 3, np.uint8)
        px[0] = data[0]
        return ImageData.rgb_slice(px, 16, 8)

    return enc, dec


def test_evaluate_image_batches_once_and_keeps_order(tmp_path):
    fake = FakeMetrics()
    cfg = EvalConfig.builder().report_dir(tmp_path).quality_levels([50, 80, 95]).metrics(MetricConfig.all()).build()
    s = EvalSession(cfg, metrics=fake)
    e1, d1 = _codec(0)
    e2, d2 = _codec(100)
    s.add_codec_with_decode("a", "1.0", e1, d1).add_codec("enc-only", "0.1", e2).add_codec_with_decode("b", "2.0", e2, d2)
    assert s.codec_count() == 3
    rep = s.evaluate_image("img", _img())
    # one batched metric call for the whole image, every pair sharing ONE reference buffer
    assert len(fake.calls) == 1 and fake.calls[0][0] == 6 and len(fake.calls[0][1]) == 1
    # results in the reference's nested-loop order: codecs outer, quality levels inner (session.rs:375-376)
    assert [(r.codec_id, r.quality) for r in rep.results] == [("a", 50.0), ("a", 80.0), ("a", 95.0), ("enc-only", 50.0),
                                                               ("enc-only", 80.0), ("enc-only", 95.0), ("b", 50.0), ("b", 80.0),
                                                               ("b", 95.0)]
    for r in rep.results:
        if r.codec_id == "enc-only":            # no decoder: default metrics, no perception (session.rs:412-429)
            assert r.metrics == MetricResult() and r.perception is None and r.decode_time is None
        else:
            assert r.metrics.ssimulacra2 == r.quality and r.perception == PerceptionLevel.from_dssim(r.metrics.dssim)
            assert r.metrics.sse is None
        assert r.bits_per_pixel == r.file_size * 8 / (16 * 8)
    assert rep.uncompressed_size == 16 * 8 * 3 and rep.best_at_size(10 ** 9).quality == 95.0
    assert rep.smallest_at_quality(0.0003).file_size == min(r.file_size for r in rep.results if r.metrics.dssim is not None and r.metrics.dssim <= 0.0003)


def test_report_formats(tmp_path):
    fake = FakeMetrics()
    cfg = EvalConfig.builder().report_dir(tmp_path / "out").quality_levels([80, 85.5]).build()
    s = EvalSession(cfg, metrics=fake)
    e, d = _codec(0)
    s.add_codec_with_decode("mozjpeg", "4.0", e, d).add_codec("x", "1", e)
    corpus = s.evaluate_corpus("run1", [("one", _img()), ("two", _img())])
    assert corpus.total_results() == 8 and corpus.codec_ids() == ["mozjpeg", "x"]
    s.write_image_report(corpus.images[0])
    s.write_corpus_report(corpus)
    j = json.load(open(tmp_path / "out" / "run1.json"))
    r0 = j["images"][0]["results"][0]
    assert set(r0) == {"codec_id", "codec_version", "quality", "file_size", "bits_per_pixel", "encode_time", "decode_time",
                       "metrics", "perception", "cached_path", "codec_params"}                       # report.rs:16-52
    assert set(r0["metrics"]) == {"dssim", "ssimulacra2", "butteraugli", "psnr"} and isinstance(r0["encode_time"], int)
    assert r0["perception"] in ("Imperceptible", "Marginal", "Subtle", "Noticeable", "Degraded")
    assert j["images"][0]["results"][2]["metrics"]["dssim"] is None and j["images"][0]["results"][2]["perception"] is None
    rows = list(csv.reader(open(tmp_path / "out" / "run1.csv")))
    assert rows[0] == ["image", "codec", "version", "quality", "file_size", "bpp", "encode_ms", "decode_ms", "dssim", "ssimulacra2",
                       "butteraugli", "psnr", "perception"]                                         # session.rs:530-544
    assert rows[1][:4] == ["one", "mozjpeg", "4.0", "80"] and rows[2][3] == "85.5"                 # f64::to_string
    assert rows[1][5] == f"{(10 + 80) * 8 / 128:.4f}" and rows[1][8] == "0.000200" and rows[1][9] == "80.00"
    assert rows[1][10] == "1.0000" and rows[1][11] == "36.00" and rows[1][12] == "IMP"
    assert rows[3][7:] == ["", "", "", "", "", ""]                                                  # encoder-only codec
    assert os.path.exists(tmp_path / "out" / "one.json")


def test_image_data_and_config_defaults():
    rgba = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    d = ImageData.rgba8(rgba)
    assert (d.width(), d.height()) == (3, 2)
    assert np.array_equal(d.to_rgb8_vec(), rgba[:, :, :3].reshape(-1))                               # alpha dropped, session.rs:100-114
    # an attached profile is applied on the device (tests/test_icc.py); untagged data passes through untouched
    assert np.array_equal(ImageData.rgb_slice(np.arange(12, dtype=np.uint8), 2, 2).to_rgb8_srgb(), np.arange(12, dtype=np.uint8))
    with pytest.raises(ValueError):
        ImageData.rgb_slice(np.zeros(11, np.uint8), 2, 2)
    c = EvalConfig.builder().report_dir("/tmp/x").build()
    assert c.quality_levels == [50.0, 60.0, 70.0, 80.0, 85.0, 90.0, 95.0] and c.metrics == MetricConfig.all()  # session.rs:273-275
    with pytest.raises(AssertionError):
        EvalConfig.builder().build()
    assert EncodeRequest(80).with_param("subsampling", "420").params == {"subsampling": "420"}
    rep = ImageReport("test.png", 1920, 1080)                                                        # report.rs:253-259
    assert rep.uncompressed_size == 1920 * 1080 * 3
    cr = CodecResult("t", "1", 80.0, 1000, 0.5, 0.1, None, MetricResult(), None)
    assert abs(cr.compression_ratio(10000) - 10.0) < 1e-3                                            # report.rs:262-279
    assert CorpusReport("c").codec_ids() == []
    s = EvalSession(c, metrics=object())                                                             # session.rs:630-637
    assert s.codec_count() == 0
    s.add_codec("test", "1.0", lambda image, req: bytes(100))
    assert s.codec_count() == 1
    assert (rep.name, rep.width, rep.height) == ("test.png", 1920, 1080)


@pytest.mark.gpu
def test_session_on_gpu_matches_direct_calls(gpu, tmp_path):
    """helpers.rs:338-383 rows + evaluate_image through the real CUDA path."""
    from codec_eval_b200.session import assert_perception_level, assert_quality, evaluate_single
    from codec_eval_b200.synth import G, cheap_distort

    w, h = 96, 64
    ref = G(3, w, h)

    def enc(image, req):
        return cheap_distort(image.to_rgb8_vec().reshape(h, w, 3), int(req.quality), seed=1).tobytes()

    def dec(data):
        return ImageData.rgb_slice(np.frombuffer(data, np.uint8), w, h)

    s = EvalSession(EvalConfig.builder().report_dir(tmp_path).quality_levels([40, 70, 95]).build(), metrics=gpu)
    s.add_codec_with_decode("blocky", "1", enc, dec)
    rep = s.evaluate_image("g3", ImageData.rgb8(ref))
    for r in rep.results:
        d = np.frombuffer(enc(ImageData.rgb8(ref), EncodeRequest(r.quality)), np.uint8)
        assert r.metrics.ssimulacra2 == gpu.calculate_ssimulacra2(ref, d, w, h)
        assert r.metrics.dssim == gpu.calculate_dssim_rgb8(ref, d, w, h)
        assert r.metrics.butteraugli == gpu.calculate_butteraugli(ref, d, w, h)
        assert r.metrics.psnr == gpu.calculate_psnr(ref, d, w, h)
        assert r.perception == PerceptionLevel.from_dssim(r.metrics.dssim)
    assert rep.results[0].metrics.ssimulacra2 < rep.results[1].metrics.ssimulacra2 < rep.results[2].metrics.ssimulacra2
    # helpers.rs tests: identical images, size mismatch, thresholds
    pat = ((np.arange(64 * 64 * 3)) % 256).astype(np.uint8).reshape(64, 64, 3)
    m = evaluate_single(pat, pat, MetricConfig.all(), gpu)
    assert m.dssim < 1e-4 and m.ssimulacra2 > 99 and m.butteraugli < 0.1
    from codec_eval_b200.metrics import DimensionMismatch

    with pytest.raises(DimensionMismatch):
        evaluate_single(pat, pat[:32], MetricConfig.all(), gpu)
    assert_quality(pat, pat, 90.0, 0.001, gpu)
    bad = ((pat.astype(int) + 50) % 256).astype(np.uint8)
    with pytest.raises(QualityBelowThreshold):
        assert_quality(pat, bad, 99.0, None, gpu)
    assert_perception_level(pat, pat, PerceptionLevel.Imperceptible, gpu)
    with pytest.raises(QualityBelowThreshold):
        assert_perception_level(pat, bad, PerceptionLevel.Imperceptible, gpu)


def test_identical_images_serialise_psnr_as_null_like_serde_json():
    """PSNR of a lossless result is +inf (src/metrics/mod.rs:326-328).  serde_json writes non-finite f64 as null; a bare
    `Infinity` token is not JSON and the reference's reader rejects it."""
    import json
    import math

    from codec_eval_b200.metrics import MetricResult
    from codec_eval_b200.session import CodecResult

    r = CodecResult("png", "1.0", 100.0, 1234, 8.0, 0.01, None,
                    MetricResult(dssim=0.0, ssimulacra2=100.0, butteraugli=0.0, psnr=math.inf), None)
    text = json.dumps(r.to_json(), allow_nan=False)
    assert "Infinity" not in text and json.loads(text)["metrics"] == {"dssim": 0.0, "ssimulacra2": 100.0, "butteraugli": 0.0, "psnr": None}
