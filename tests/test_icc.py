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
