"""ICC matrix/TRC -> sRGB on the device (src/metrics/icc.rs:69-103).

CPU: the numpy oracle against a real CMS (Little CMS 2 via Pillow's ImageCms) within 2 code values -- the reference's
moxcms is not available, so parity against it is unpinned -- plus the reference's own ICC tests (icc.rs:137-160:
sRGB passthrough, from_icc_bytes rules).  GPU: the CUDA transform against the oracle, bit for bit."""
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import icc_profiles as P  # noqa: E402

from codec_eval_b200.synth import G  # noqa: E402
from oracle import icc_oracle as IO  # noqa: E402

PROFILES = {"display_p3": P.display_p3, "adobe_rgb": P.adobe_rgb, "rec2020_table": P.rec2020_table, "srgb_like": P.srgb_like}


def _lcms(rgb, icc):
    from PIL import Image, ImageCms

    src = ImageCms.ImageCmsProfile(io.BytesIO(icc))
    dst = ImageCms.createProfile("sRGB")
    im = Image.fromarray(rgb, "RGB")
    out = ImageCms.profileToProfile(im, src, dst, renderingIntent=ImageCms.Intent.RELATIVE_COLORIMETRIC, outputMode="RGB")
    return np.asarray(out)


def _samples():
    ramp = np.stack(np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 5), np.arange(0, 256, 51), indexing="ij"), -1)
    return np.concatenate([G(7, 96, 64).reshape(-1, 3), ramp.reshape(-1, 3).astype(np.uint8)]).reshape(-1, 1, 3).copy()


@pytest.mark.parametrize("name", sorted(PROFILES))
def test_oracle_close_to_lcms2(name):
    icc = PROFILES[name]()
    px = _samples()
    got, exp = IO.transform_to_srgb(px, icc), _lcms(px, icc)
    d = np.abs(got.astype(int) - exp.astype(int))
    assert d.max() <= 2, (name, d.max())
    assert d.mean() < 0.35, (name, d.mean())


def test_srgb_like_profile_is_nearly_identity():
    px = _samples()
    d = np.abs(IO.transform_to_srgb(px, P.srgb_like()).astype(int) - px.astype(int))
    assert d.max() <= 1


def test_reference_rows_passthrough_and_profile_rules():
    rgb = np.array([100, 150, 200, 50, 100, 150], np.uint8)           # icc.rs:141-146 test_srgb_passthrough
    assert np.array_equal(IO.transform_to_srgb(rgb, None), rgb)
    assert np.array_equal(IO.transform_to_srgb(rgb, b""), rgb)        # from_icc_bytes: empty => Srgb (icc.rs:49-54)


def test_unusable_profiles_are_refused():
    for bad in (b"", b"\0" * 200, P.display_p3()[:100],
                P.make_profile(P.P3, P.D65, P.SRGB_PARA, space=b"GRAY"),
                P.make_profile(P.P3, P.D65, P.SRGB_PARA, pcs=b"Lab "),
                P.make_profile(P.P3, P.D65, P.SRGB_PARA, drop=(b"gXYZ",)),
                P.make_profile(P.P3, P.D65, P.SRGB_PARA, drop=(b"bTRC",))):
        if not bad:
            continue
        with pytest.raises(IO.IccError):
            IO.tables(bad)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PROFILES))
def test_cuda_matches_oracle_bit_for_bit(gpu, name):
    icc = PROFILES[name]()
    for (w, h) in [(1, 1), (13, 7), (96, 64), (768, 512)]:
        img = G(w + h, w, h)
        got = gpu.transform_to_srgb(img, w, h, icc).reshape(h, w, 3)
        exp = IO.transform_to_srgb(img, icc)
        assert np.array_equal(got, exp), (name, w, h, int(np.abs(got.astype(int) - exp).max()))
    px = _samples()
    n = px.shape[0]
    assert np.array_equal(gpu.transform_to_srgb(px, n, 1, icc).reshape(n, 1, 3), IO.transform_to_srgb(px, icc))


@pytest.mark.gpu
def test_cuda_close_to_lcms2(gpu):
    icc = P.display_p3()
    img = G(3, 160, 96)
    d = np.abs(gpu.transform_to_srgb(img, 160, 96, icc).reshape(96, 160, 3).astype(int) - _lcms(img, icc).astype(int))
    assert d.max() <= 2


@pytest.mark.gpu
def test_cuda_passthrough_and_errors(gpu):
    from codec_eval_b200.metrics import MetricCalculation

    rgb = np.array([100, 150, 200, 50, 100, 150], np.uint8)
    assert np.array_equal(gpu.transform_to_srgb(rgb, 2, 1, None), rgb)            # icc.rs:141-146
    assert np.array_equal(gpu.transform_to_srgb(rgb, 2, 1, b""), rgb)
    for bad in (b"\0" * 200, P.make_profile(P.P3, P.D65, P.SRGB_PARA, pcs=b"Lab "),
                P.make_profile(P.P3, P.D65, P.SRGB_PARA, drop=(b"gXYZ",))):
        with pytest.raises(MetricCalculation) as e:
            gpu.transform_to_srgb(rgb, 2, 1, bad)
        assert e.value.metric == "ICC"
    with pytest.raises(AssertionError):
        gpu.transform_to_srgb(rgb, 3, 1, None)


@pytest.mark.gpu
def test_image_data_to_rgb8_srgb_applies_the_profile(gpu):
    from codec_eval_b200.session import ImageData

    img = G(9, 64, 48)
    icc = P.adobe_rgb()
    plain = ImageData.rgb_slice(img.reshape(-1), 64, 48)
    tagged = ImageData.rgb_slice_with_icc(img.reshape(-1), 64, 48, icc)
    assert np.array_equal(plain.to_rgb8_srgb(gpu), img.reshape(-1))
    assert np.array_equal(tagged.to_rgb8_srgb(gpu), IO.transform_to_srgb(img, icc).reshape(-1))
    # a wide-gamut source pushed into sRGB changes the metric input: the scores differ from the untagged ones
    a = gpu.calculate_ssimulacra2(img, tagged.to_rgb8_srgb(gpu), 64, 48)
    assert a < 100.0


@pytest.mark.gpu
def test_icc_metric_variants_on_gpu(gpu):
    """calculate_*_icc == the plain metric on the transformed buffers; profiles on one or both sides; sRGB = plain."""
    from codec_eval_b200.metrics import ColorProfile, MetricCalculation
    from codec_eval_b200.synth import J

    w, h = 96, 64
    ref = G(21, w, h)
    dist = J(ref, 70, 2)
    p3, adobe = ColorProfile.Icc(P.display_p3()), ColorProfile.from_icc_bytes(P.adobe_rgb())
    srgb = ColorProfile.Srgb
    r_p3 = IO.transform_to_srgb(ref, p3.icc)
    d_ad = IO.transform_to_srgb(dist, adobe.icc)
    assert np.array_equal(gpu.transform_profile_to_srgb(ref, p3).reshape(h, w, 3), r_p3)
    pr, pt = gpu.prepare_for_comparison(ref, p3, dist, adobe)
    assert np.array_equal(pr.reshape(h, w, 3), r_p3) and np.array_equal(pt.reshape(h, w, 3), d_ad)
    assert gpu.calculate_ssimulacra2_icc(ref, srgb, dist, srgb, w, h) == gpu.calculate_ssimulacra2(ref, dist, w, h)
    assert gpu.calculate_ssimulacra2_icc(ref, p3, dist, adobe, w, h) == gpu.calculate_ssimulacra2(r_p3, d_ad, w, h)
    assert gpu.calculate_butteraugli_icc(ref, srgb, dist, adobe, w, h) == gpu.calculate_butteraugli(ref, d_ad, w, h)
    assert gpu.calculate_dssim_icc(ref, p3, dist, srgb, w, h) == gpu.calculate_dssim_rgb8(r_p3, dist, w, h)
    # same profile on both sides of identical images: still identical after the transform
    assert gpu.calculate_ssimulacra2_icc(ref, p3, ref, p3, w, h) == 100.0
    with pytest.raises(MetricCalculation) as e:
        gpu.calculate_dssim_icc(ref, ColorProfile.Icc(b"\0" * 200), dist, srgb, w, h)
    assert e.value.metric == "ICC"


# ---- the *_icc metric variants (src/metrics/{ssimulacra2.rs:135-147, butteraugli.rs:150-162, dssim.rs:158-174}) are host
# wiring over entries the GPU tests above already cover: check the wiring on the CPU with a recording stand-in for the
# C library (no compute; the product class itself still refuses to start without a GPU)
class _FakeLib:
    def __init__(self):
        self.calls = []

    def ce_transform_to_srgb(self, h, src, n, w, hgt, icc, icc_len, dst):
        import ctypes as C

        self.calls.append(("icc", n, w, hgt, bytes(icc), icc_len))
        buf = (C.c_ubyte * n).from_address(src)
        out = (C.c_ubyte * n).from_address(dst)
        for i in range(n):
            out[i] = 255 - buf[i]  # a recognisable stand-in transform
        return 0

    def _metric(self, name, r, rn, t, tn, w, h, out):
        import ctypes as C

        a = np.frombuffer((C.c_ubyte * rn).from_address(r), np.uint8).astype(np.int64)
        b = np.frombuffer((C.c_ubyte * tn).from_address(t), np.uint8).astype(np.int64)
        self.calls.append((name, w, h))
        out._obj.value = float(np.abs(a - b).sum())
        return 0

    def ce_ssimulacra2(self, h, r, rn, t, tn, w, hh, out):
        return self._metric("ssim2", r, rn, t, tn, w, hh, out)

    def ce_dssim_rgb8(self, h, r, rn, t, tn, w, hh, out):
        return self._metric("dssim", r, rn, t, tn, w, hh, out)

    def ce_butteraugli(self, h, r, rn, t, tn, w, hh, intensity, out, pn):
        return self._metric("ba", r, rn, t, tn, w, hh, out)


def _fake_ctx():
    from codec_eval_b200.metrics import GpuMetrics

    ctx = object.__new__(GpuMetrics)
    ctx._L, ctx._h = _FakeLib(), None
    return ctx


def test_color_profile_mirrors_the_reference_enum():
    from codec_eval_b200.metrics import ColorProfile

    assert ColorProfile.Srgb.is_srgb() and ColorProfile().is_srgb()
    assert not ColorProfile.Icc(b"\x01\x02").is_srgb()
    # icc.rs:49-54 and its tests (:139-154): None and empty bytes mean sRGB
    assert ColorProfile.from_icc_bytes(None).is_srgb() and ColorProfile.from_icc_bytes(b"").is_srgb()
    assert not ColorProfile.from_icc_bytes(b"\x00" * 4).is_srgb()


def test_icc_variants_transform_each_side_then_call_the_plain_metric():
    from codec_eval_b200.metrics import ColorProfile, MetricCalculation

    ctx = _fake_ctx()
    rng = np.random.default_rng(5)
    w, h = 6, 4
    ref = rng.integers(0, 256, w * h * 3, dtype=np.uint8)
    tst = rng.integers(0, 256, w * h * 3, dtype=np.uint8)
    prof = ColorProfile.Icc(b"PROFILE")

    # sRGB on both sides: no transform call, buffers passed through (icc.rs:73)
    r, t = ctx.prepare_for_comparison(ref, ColorProfile.Srgb, tst, ColorProfile.Srgb)
    assert np.array_equal(r, ref) and np.array_equal(t, tst) and r is not ref and ctx._L.calls == []

    # only the side that carries a profile is transformed, as one row of len/3 pixels
    want = float(np.abs(ref.astype(np.int64) - (255 - tst.astype(np.int64))).sum())
    for name, fn in (("ssim2", ctx.calculate_ssimulacra2_icc), ("ba", ctx.calculate_butteraugli_icc),
                     ("dssim", ctx.calculate_dssim_icc)):
        ctx._L.calls.clear()
        got = fn(ref, ColorProfile.Srgb, tst, prof, w, h)
        assert got == want, name
        assert ctx._L.calls == [("icc", ref.size, w * h, 1, b"PROFILE", 7), (name, w, h)]
    ctx._L.calls.clear()
    ctx.calculate_ssimulacra2_icc(ref, prof, tst, prof, w, h)
    assert [c[0] for c in ctx._L.calls] == ["icc", "icc", "ssim2"]  # reference first (icc.rs:127-128)

    # an empty Icc(..) and a ragged buffer are ICC errors, raised before the metric's own validation
    for bad_args in ((ref, ColorProfile.Icc(b""), tst, ColorProfile.Srgb), (ref[:-1], prof, tst, ColorProfile.Srgb)):
        ctx._L.calls.clear()
        with pytest.raises(MetricCalculation) as e:
            ctx.calculate_butteraugli_icc(*bad_args, w, h)
        assert e.value.metric == "ICC" and ctx._L.calls == []
