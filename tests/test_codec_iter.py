"""codec-iter twin: PPM source cache, EvalPoint / Baseline JSON, compare_with_baseline, and the
batched run_eval loops.  Format tests run on the CPU; the eval loops need the GPU."""
import io
import json
import os

import numpy as np
import pytest

from codec_eval_b200.synth import G, J

HERE = os.path.dirname(os.path.abspath(__file__))


def _ci():
    from codec_eval_b200 import codec_iter

    return codec_iter


def test_ppm_round_trip_and_header_variants():
    ci = _ci()
    img = G(1, 37, 21)
    data = ci.encode_ppm(img)
    assert data.startswith(b"P6\n37 21\n255\n") and len(data) == 13 + 37 * 21 * 3
    assert np.array_equal(ci.decode_ppm(data), img)
    odd = b"P6 # made by hand\n# another comment\n37\t21\n255\n" + img.tobytes()
    assert np.array_equal(ci.decode_ppm(odd), img)
    with pytest.raises(ValueError):
        ci.decode_ppm(b"P5\n1 1\n255\n\x00")
    with pytest.raises(ValueError):
        ci.decode_ppm(b"P6\n4 4\n255\n\x00\x00")
    with pytest.raises(ValueError):
        ci.decode_ppm(b"P6\n1 1\n65535\n\x00\x00\x00\x00\x00\x00")


def test_source_cache_layout(tmp_path):
    """source.rs:62-92: <corpus>/.codec-iter-cache/<stem>.ppm is read first; a PNG is decoded once and cached."""
    from PIL import Image

    ci = _ci()
    corpus = str(tmp_path)
    a, b = G(2, 24, 16), G(3, 16, 24)
    Image.fromarray(a, "RGB").save(os.path.join(corpus, "a.png"))
    os.makedirs(os.path.join(corpus, ci.CACHE_DIR))
    open(os.path.join(corpus, ci.CACHE_DIR, "b.ppm"), "wb").write(ci.encode_ppm(b))   # cache only, no PNG
    imgs = ci.load_by_names(corpus, ["a.png", "b.png"])
    assert [(i.name, i.width, i.height) for i in imgs] == [("a.png", 24, 16), ("b.png", 16, 24)]
    assert np.array_equal(imgs[0].pixels, a) and np.array_equal(imgs[1].pixels, b)
    assert os.path.exists(os.path.join(corpus, ci.CACHE_DIR, "a.ppm"))               # written on first load
    os.remove(os.path.join(corpus, "a.png"))
    assert np.array_equal(ci.load_by_names(corpus, ["a.png"])[0].pixels, a)          # now served from the cache
    with pytest.raises(FileNotFoundError):
        ci.load_by_names(corpus, ["missing.png"])
    assert [i.name for i in ci.load_all_from_dir(corpus)] == []                       # only *.png files are listed


def test_baseline_json_is_byte_identical_to_the_reference_writer(tmp_path):
    ci = _ci()
    text = open(os.path.join(HERE, "golden", "codec_iter_baseline_excerpt.json")).read()
    b = ci.Baseline.from_json(text)
    assert b.format == "jpeg" and b.config_summary == "zenjpeg-420-ycbcr-prog" and len(b.points) == 6
    assert b.points[0] == ci.EvalPoint("pexels-photo-951408.png", 50, 0.72332763671875, 67.06036004649532, 23702, 3)
    assert b.to_json() == text                       # serde_json::to_string_pretty layout, field order and float text
    ci.save_baseline(str(tmp_path), b)
    assert ci.baseline_path(str(tmp_path), "jpeg").endswith("jpeg.json")
    again = ci.load_baseline(str(tmp_path), "jpeg")
    assert again.points == b.points and again.created_at == b.created_at
    assert ci.load_baseline(str(tmp_path), "webp") is None


def test_compare_with_baseline_rows():
    ci = _ci()
    base = ci.Baseline("jpeg", "x", "/c", points=[ci.EvalPoint("a", 50, 1.0, 60.0, 1, 0), ci.EvalPoint("b", 50, 2.0, 70.0, 1, 0),
                                                  ci.EvalPoint("a", 90, 3.0, 90.0, 1, 0)])
    cur = [ci.EvalPoint("a", 50, 1.0, 62.0, 1, 0), ci.EvalPoint("b", 50, 1.0, 70.0, 1, 0), ci.EvalPoint("a", 75, 2.0, 80.0, 1, 0)]
    assert ci.aggregate_by_quality(cur) == {50: (1.0, 66.0), 75: (2.0, 80.0)}
    rows = ci.compare_with_baseline(cur, base)
    assert [r.quality for r in rows] == [50, 75]
    r = rows[0]
    assert (r.bpp, r.ssim2, r.delta_bpp, r.delta_ssim2) == (1.0, 66.0, -0.5, 1.0) and r.pareto == 1.0 - (-0.5) * 10.0
    assert (rows[1].delta_bpp, rows[1].delta_ssim2, rows[1].pareto) == (0.0, 0.0, 0.0)   # quality absent from the baseline


# ------------------------------------------------------------------ GPU
def _pillow_codec(ci, ss=2):
    from PIL import Image

    def enc(px, q):
        buf = io.BytesIO()
        Image.fromarray(px, "RGB").save(buf, format="JPEG", quality=q, subsampling=ss)
        return buf.getvalue()

    def dec(data):
        return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))

    return ci.CodecConfig(enc, dec, f"pillow-jpeg-ss{ss}")


@pytest.mark.gpu
def test_run_eval_matches_per_pair_calls(gpu):
    ci = _ci()
    images = [ci.SourceImage(f"img{i}.png", 96, 64, G(20 + i, 96, 64)) for i in range(3)]
    codec = _pillow_codec(ci)
    res = ci.run_eval(images, [50, 75, 95], codec, metrics=gpu)
    assert res.config_summary == "pillow-jpeg-ss2" and len(res.points) == 9
    assert [(p.image, p.quality) for p in res.points[:4]] == [("img0.png", 50), ("img0.png", 75), ("img0.png", 95), ("img1.png", 50)]
    for p in res.points:
        im = images[int(p.image[3])]
        enc = codec.encode(im.pixels, p.quality)
        assert p.size_bytes == len(enc) and p.bpp == len(enc) * 8.0 / (96 * 64)
        assert p.ssim2 == gpu.calculate_ssimulacra2(im.pixels, codec.decode(enc), 96, 64)
    # higher quality => higher score and more bits, per image (the sweep's sanity rule)
    for i in range(3):
        s = [p for p in res.points if p.image == f"img{i}.png"]
        assert s[0].ssim2 < s[1].ssim2 < s[2].ssim2 and s[0].bpp < s[1].bpp < s[2].bpp


@pytest.mark.gpu
def test_on_device_sweep_equals_the_real_codec_sweep(gpu, tmp_path):
    """Decoded images generated on the device are bit-identical to libjpeg-turbo's, so the scores are identical."""
    ci = _ci()
    images = [ci.SourceImage(f"s{i}.png", 128, 96, G(40 + i, 128, 96)) for i in range(2)] + \
             [ci.SourceImage("t.png", 77, 35, G(50, 77, 35))]
    codec = _pillow_codec(ci)
    qs = [60, 85]
    host = ci.run_eval(images, qs, codec, metrics=gpu)
    dev = ci.run_eval_jpeg_sweep(images, qs, 2, size_of=lambda px, q: len(codec.encode(px, q)), metrics=gpu)
    assert dev.config_summary == "ce-gpu-jpeg-420-ycbcr-baseline"
    assert [(p.image, p.quality, p.ssim2, p.size_bytes, p.bpp) for p in dev.points] == \
           [(p.image, p.quality, p.ssim2, p.size_bytes, p.bpp) for p in host.points]
    b = ci.Baseline("jpeg", dev.config_summary, "/synthetic", points=dev.points)
    ci.save_baseline(str(tmp_path), b)
    rows = ci.compare_with_baseline(host.points, ci.load_baseline(str(tmp_path), "jpeg"))
    assert all(r.delta_bpp == 0.0 and r.delta_ssim2 == 0.0 for r in rows)
    json.loads(b.to_json())


def test_ppm_decoder_on_arbitrary_images_and_damaged_files():
    """encode -> decode is the identity for any 8-bit RGB image; a damaged file raises ValueError, never anything else."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    ci = _ci()

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 17), st.integers(1, 13), st.integers(0, 2**32 - 1))
    def roundtrip(w, h, seed):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        data = ci.encode_ppm(img)
        assert np.array_equal(ci.decode_ppm(data), img)
        assert np.array_equal(ci.decode_ppm(data.replace(b"P6\n", b"P6 # made by a test\n\t", 1)), img)
        for cut in (3, len(data) // 2, len(data) - 1):
            with pytest.raises(ValueError):
                ci.decode_ppm(data[:cut])

    roundtrip()
    for bad in (b"", b"P6", b"P6\n0 4\n255\n", b"P6\n-2 4\n255\n" + b"\0" * 24, b"P6\nx y\n255\n", b"P6\n2 2\n0\n" + b"\0" * 12):
        with pytest.raises(ValueError):
            ci.decode_ppm(bad)
