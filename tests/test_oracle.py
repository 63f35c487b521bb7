"""CPU tests of the oracle: the reference's own unit-test rows re-expressed
against oracle/ce_oracle.c, plus the committed golden table and structural properties."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pairs.npz")


# ---- src/metrics/mod.rs:369-383
def test_psnr_identical(O):
    data = np.full(100 * 100 * 3, 128, np.uint8)
    assert np.isinf(O.psnr(data, data, 100, 100))


def test_psnr_different(O):
    ref = np.full(100 * 100 * 3, 100, np.uint8)
    test = np.full(100 * 100 * 3, 110, np.uint8)
    p = O.psnr(ref, test, 100, 100)
    assert 28.0 < p < 29.0
    assert abs(p - 10 * np.log10(255.0 ** 2 / 100.0)) < 1e-12
    assert O.sse(ref, test) == 100 * 100 * 3 * 100
    assert O.psnr_from_sse(O.sse(ref, test), 100, 100) == p


# ---- src/metrics/ssimulacra2.rs:154-182
def test_ssim2_identical_images(O):
    ref = (np.arange(100 * 100 * 3) % 256).astype(np.uint8)
    assert O.ssimulacra2(ref, ref, 100, 100) > 99.0
    assert O.ssimulacra2(ref, ref, 100, 100) == 100.0


def test_ssim2_different_images(O):
    ref = np.full(100 * 100 * 3, 100, np.uint8)
    test = np.full(100 * 100 * 3, 200, np.uint8)
    assert O.ssimulacra2(ref, test, 100, 100) < 80.0


def test_ssim2_too_small(O):
    a = np.zeros(7 * 7 * 3, np.uint8)
    with pytest.raises(O.OracleError):
        O.ssimulacra2(a, a, 7, 7)


def test_ssim2_scale_count(O):
    # upstream checks the size before halving: 64x64 and 100x100 run 5 scales, >= 256 min side runs 6
    for (w, h, ns) in [(64, 64, 5), (100, 100, 5), (256, 256, 6), (8, 8, 2), (40, 24, 3)]:
        a = (np.arange(w * h * 3) % 251).astype(np.uint8)
        _, sums = O.ssimulacra2_ex(a, a, w, h)
        assert sums.shape[0] == ns, (w, h, sums.shape)


# ---- src/metrics/butteraugli.rs:169-207
def test_butteraugli_identical_different_intensity(O):
    ref = (np.arange(100 * 100 * 3) % 256).astype(np.uint8)
    assert O.butteraugli(ref, ref, 100, 100)[0] < 0.01
    a = np.full(100 * 100 * 3, 100, np.uint8)
    b = np.full(100 * 100 * 3, 200, np.uint8)
    assert O.butteraugli(a, b, 100, 100)[0] > 1.0
    assert O.butteraugli(ref, ref, 100, 100, 250.0)[0] < 0.01


# ---- src/metrics/dssim.rs:181-273
def test_dssim_identical_and_different(O):
    a = np.full((100, 100, 4), 0.5, np.float32)
    a[..., 3] = 1.0
    assert O.dssim_rgbaf32(a, a, 100, 100) < 1e-4
    b = np.full((100, 100, 4), 0.3, np.float32)
    c = np.full((100, 100, 4), 0.7, np.float32)
    b[..., 3] = c[..., 3] = 1.0
    assert O.dssim_rgbaf32(b, c, 100, 100) > 0.0


def test_rgb8_rgba8_conversion(O):
    img = O.rgb8_to_dssim_image(np.array([255, 255, 255, 0, 0, 0], np.uint8), 2, 1)
    assert abs(img[0, 0, 0] - 1.0) < 1e-3 and abs(img[0, 0, 3] - 1.0) < 1e-3 and abs(img[0, 1, 0]) < 1e-3
    img = O.rgba8_to_dssim_image(np.array([255, 255, 255, 255, 0, 0, 0, 128], np.uint8), 2, 1)
    assert abs(img[0, 0, 3] - 1.0) < 1e-3 and abs(img[0, 1, 3] - 0.502) < 0.01


def test_dssim_rgb8_equals_rgbaf32(O):
    z = np.load(GOLD)
    ref, dist = z["ref0"], z["dist0"]
    a = O.dssim(ref, dist, 64, 64)
    b = O.dssim_rgbaf32(O.rgb8_to_dssim_image(ref, 64, 64), O.rgb8_to_dssim_image(dist, 64, 64), 64, 64)
    assert a == b


# ---- src/metrics/xyb.rs:260-301
def test_xyb_roundtrip(O):
    rgb = (np.arange(64 * 64 * 3) % 256).astype(np.uint8)
    assert O.xyb_roundtrip(rgb, 64, 64).size == rgb.size
    rgb = ((np.arange(32 * 32 * 3) * 7) % 256).astype(np.uint8)
    assert np.array_equal(O.xyb_roundtrip(rgb, 32, 32), O.xyb_roundtrip(rgb, 32, 32))
    g = np.arange(0, 256, 16, dtype=np.uint8)
    cube = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    out = O.xyb_roundtrip(cube, cube.shape[0], 1).reshape(-1, 3)
    md = np.abs(out.astype(int) - cube.astype(int)).max()
    assert 0 < md <= 30


# ---- golden table (regression pin of the oracle) and structural checks
def test_golden_table(O):
    z = np.load(GOLD)
    for k, row in enumerate(z["table"]):
        seed, w, h, q, ss, sse, ps, s2, ds, ba, pn = row
        w, h = int(w), int(h)
        ref, dist = z[f"ref{k}"], z[f"dist{k}"]
        assert O.sse(ref, dist) == int(sse)
        assert O.psnr(ref, dist, w, h) == ps
        assert abs(O.ssimulacra2(ref, dist, w, h) - s2) < 1e-9
        assert abs(O.dssim(ref, dist, w, h) - ds) < 1e-12
        m, p = O.butteraugli(ref, dist, w, h)
        assert abs(m - ba) < 1e-6 and abs(p - pn) < 1e-9
    assert np.array_equal(O.xyb_roundtrip(z["ref0"], 64, 64).reshape(64, 64, 3), z["xyb_rt0"])


def test_probe_anchors(O):
    """an independent numpy restatement produced these levels on G(0,256,256) + Pillow JPEG.
    DSSIM / Butteraugli agree to the printed digits; SSIMULACRA2 within 0.06 (fp32 recurrence vs FIR)."""
    from codec_eval_b200.synth import G, J

    ref = G(0, 256, 256)
    anchors = {30: (64.920, 0.003860, 2.9728, 1.4364), 80: (78.987, 0.001286, 1.6587, 0.9388), 95: (88.249, 0.000400, 1.1509, 0.5905)}
    for q, (s2, ds, ba, pn) in anchors.items():
        d = J(ref, q, 2)
        assert abs(O.ssimulacra2(ref, d, 256, 256) - s2) < 0.06
        assert abs(O.dssim(ref, d, 256, 256) - ds) < 2e-6
        m, p = O.butteraugli(ref, d, 256, 256)
        assert abs(m - ba) < 2e-4 and abs(p - pn) < 2e-4


def test_monotone_in_quality(O):
    from codec_eval_b200.synth import G, J

    ref = G(7, 128, 96)
    prev = None
    for q in (30, 50, 70, 90):
        d = J(ref, q, 2)
        cur = (O.ssimulacra2(ref, d, 128, 96), -O.dssim(ref, d, 128, 96), -O.butteraugli(ref, d, 128, 96)[1])
        if prev is not None:
            assert all(c > p for c, p in zip(cur, prev)), (q, cur, prev)
        prev = cur


def test_rgauss_is_9tap_fir(O):
    """The sigma-1.5 recursive Gaussian equals a 9-tap symmetric FIR in exact arithmetic."""
    imp = np.zeros((1, 41), np.float32)
    imp[0, 20] = 1.0
    taps = np.array([0.00941436781, 0.03601111466, 0.1093353728, 0.212928592, 0.2646211055])
    # the vertical pass over a 1-row image multiplies by the centre tap
    hrow = O.rgauss_blur(imp)[0] / taps[4]
    full = np.concatenate([taps, taps[-2::-1]])
    assert np.allclose(hrow[16:25], full, atol=2e-6)
    assert np.abs(hrow[:16]).max() < 2e-6 and np.abs(hrow[25:]).max() < 2e-6


def test_fast_log2_and_cbrt_accuracy(O):
    xs = np.linspace(0.004, 1.2, 20001, dtype=np.float32)
    c = np.array([O.lib().ceo_cbrtf(float(x)) for x in xs], np.float32)
    assert np.array_equal(c, np.cbrt(xs.astype(np.float64)).astype(np.float32))     # correctly rounded, like yuvxyb-math's
    ys = np.geomspace(1e-3, 1e4, 2000).astype(np.float32)
    l2 = np.array([O.lib().ceo_ba_fast_log2f(float(y)) for y in ys])
    assert np.abs(l2 - np.log2(ys.astype(np.float64))).max() < 2e-5


def test_batch_matches_single(O):
    z = np.load(GOLD)
    refs = np.stack([z["ref0"], z["ref0"]])
    dists = np.stack([z["dist0"], z["ref0"]])
    res = O.evaluate_batch(refs, dists, 64, 64, 15, threads=2)
    assert res[0].sse == int(z["table"][0][5]) and res[1].sse == 0
    assert abs(res[0].ssimulacra2 - z["table"][0][7]) < 1e-9
    assert res[1].ssimulacra2 == 100.0 and res[1].dssim == 0.0 and res[1].butteraugli == 0.0 and np.isinf(res[1].psnr)


def test_two_restatements_agree(O):
    """the C oracle and an independent numpy restatement written from Appendix A.3 / A.4
    (oracle/np_restatement.py) must give the same numbers.  DSSIM agrees to the last bit of the pooled sums;
    SSIMULACRA2 to 1e-9 -- both now take the cube root the reference's dependency takes (yuvxyb-math's cbrtf, which is
    the correctly rounded one: the C oracle restates its double-precision Halley steps, numpy rounds np.cbrt in float64
    once).  Round 1's own 0.77-ulp fp32 cbrt moved the score by 0.0003 ... 0.016 against this one, more than the 0.01
    contract at 768x512 / q90, which is why it was replaced."""
    from codec_eval_b200.synth import G, J
    from oracle import np_restatement as N

    cases = [(64, 48, 60), (160, 96, 85), (100, 100, 40), (77, 35, 70), (9, 33, 50)]
    for w, h, q in cases:
        r = G(w + h, w, h)
        d = J(r, q, 2)
        e_ds, e_s2 = O.dssim(r, d, w, h), O.ssimulacra2(r, d, w, h)
        assert abs(N.dssim(r, d) - e_ds) <= 1e-9 * e_ds, (w, h, q)
        assert abs(N.ssimulacra2(r, d) - e_s2) < 1e-9, (w, h, q)
    same = G(1, 40, 40)
    assert N.ssimulacra2(same, same) == 100.0 and N.dssim(same, same) == 0.0
    # Butteraugli (Appendix A.5): float64-accumulated blurs and an exact log in numpy vs fp32 fused chains and
    # FastLog2f in C => agreement to ~1e-6 relative; a wrong constant or tap would show at the percent level
    for w, h, q in cases + [(256, 256, 80)]:
        r = G(w + h, w, h)
        d = J(r, q, 2)
        mx, pn = N.butteraugli(r, d)
        emx, epn = O.butteraugli(r, d, w, h)
        assert abs(mx - emx) <= 2e-5 * emx and abs(pn - epn) <= 2e-5 * epn, (w, h, q, mx, emx, pn, epn)
    assert N.butteraugli(same, same) == (0.0, 0.0)


def _fnv1a64(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def test_reference_scores_pin_the_oracle(O):
    """The pin that is still missing: tests/golden/reference_scores.json holds the outputs of the
    real codec-eval functions for the golden pairs, produced by rust/pin-parity on a machine with a Rust toolchain.
    While that file is absent parity with the three crates is UNPINNED and this test skips; once it is committed the
    oracle must meet the contract tolerances (and the GPU tests, which compare against the oracle, inherit the pin)."""
    import json

    path = os.path.join(os.path.dirname(GOLD), "reference_scores.json")
    if not os.path.exists(path):
        pytest.skip("parity unpinned: no tests/golden/reference_scores.json (run rust/pin-parity to create it)")
    z = np.load(GOLD)
    for row in json.load(open(path))["cases"]:
        k, w, h = row["case"], row["width"], row["height"]
        r, d = z[f"ref{k}"], z[f"dist{k}"]
        assert r.shape == (h, w, 3)
        assert O.psnr(r, d, w, h) == pytest.approx(row["psnr"], rel=1e-12)
        assert abs(O.ssimulacra2(r, d, w, h) - row["ssimulacra2"]) <= 0.01
        assert abs(O.dssim(r, d, w, h) - row["dssim"]) <= 1e-4 * row["dssim"]
        assert abs(O.butteraugli(r, d, w, h)[0] - row["butteraugli"]) <= 1e-3 * row["butteraugli"]
        assert abs(O.butteraugli(r, d, w, h, 250.0)[0] - row["butteraugli_250"]) <= 1e-3 * row["butteraugli_250"]
        rt = O.xyb_roundtrip(r, w, h)
        assert "%016x" % _fnv1a64(np.asarray(rt, np.uint8).tobytes()) == row["xyb_roundtrip_fnv1a64"]
        assert abs(O.ssimulacra2(np.asarray(rt).reshape(h, w, 3), d, w, h) - row["ssimulacra2_xyb_ref"]) <= 0.01


def test_plausibility_on_a_photograph(O):
    """NOT a parity pin -- a guard against gross constant errors (SURVEY A.7 grade-B items).  On a real photograph
    (scikit-learn's bundled china.jpg) compressed by libjpeg-turbo 4:2:0 the three metrics must land where the
    reference's own tables put such encodes: its PerceptionLevel boundaries (src/metrics/mod.rs:189-232) line the metrics
    up as DSSIM 0.0003 / SSIMULACRA2 90 / Butteraugli 1.0 (imperceptible) ... 0.0015 / 70 / 3.0 (subtle), and its README
    calibration (mozjpeg 4:2:0 on CID22: SSIMULACRA2 65.1 and Butteraugli 4.38 at ~0.7 bpp) ties SSIMULACRA2 ~65 to
    Butteraugli ~3-5.  Measured with this oracle: q95 87.3 / 0.00026 / 1.12, q75 70.1 / 0.0024 / 2.80, q30 42.5 / 0.0099 / 4.34."""
    sk = pytest.importorskip("sklearn.datasets")
    from codec_eval_b200.synth import J

    img = np.ascontiguousarray(sk.load_sample_image("china.jpg"))
    h, w, _ = img.shape
    bands = {  # quality: (ssimulacra2 lo, hi), (dssim lo, hi), (butteraugli lo, hi)
        95: ((80.0, 95.0), (1e-4, 8e-4), (0.5, 2.5)),
        75: ((60.0, 80.0), (1e-3, 4e-3), (1.5, 4.5)),
        30: ((25.0, 55.0), (4e-3, 2e-2), (3.0, 8.0)),
    }
    prev = None
    for q in (95, 75, 30):
        d = J(img, q, 2)
        got = (O.ssimulacra2(img, d, w, h), O.dssim(img, d, w, h), O.butteraugli(img, d, w, h)[0])
        for v, (lo, hi) in zip(got, bands[q]):
            assert lo <= v <= hi, (q, got)
        if prev:
            assert got[0] < prev[0] and got[1] > prev[1] and got[2] > prev[2]
        prev = got


def test_two_restatements_agree_at_kodak_size(O):
    """The same cross-check at a BASELINE size (768x512, JPEG q75 4:2:0 and q90 4:4:4), not only on thumbnails: the
    numpy restatement (np.cbrt in float64 rounded once, float64 blurs for Butteraugli, exact log) and the C oracle
    (yuvxyb-math's cbrtf restated, fp32 chains, FastLog2f) give the same SSIMULACRA2 and DSSIM to 1e-9 and the same
    Butteraugli to 2e-5 relative."""
    from codec_eval_b200.synth import G, J
    from oracle import np_restatement as N

    w, h = 768, 512
    r = G(3, w, h)
    for q, ss in ((75, 2), (90, 0)):
        d = J(r, q, ss)
        e_ds, e_s2 = O.dssim(r, d, w, h), O.ssimulacra2(r, d, w, h)
        emx, epn = O.butteraugli(r, d, w, h)
        assert abs(N.dssim(r, d) - e_ds) <= 1e-9 * e_ds
        assert abs(N.ssimulacra2(r, d) - e_s2) < 1e-9
        mx, pn = N.butteraugli(r, d)
        assert abs(mx - emx) <= 2e-5 * emx and abs(pn - epn) <= 2e-5 * epn


def test_xyb_roundtrip_libm_choice_moves_few_bytes(O):
    """The oracle pins xyb_roundtrip (src/metrics/xyb.rs:60-100,225-253) to correctly rounded cbrt / pow; the Rust
    reference calls the platform libm (glibc cbrtf is off by one ulp for ~11 % of inputs).  Quantified over a lattice of
    the whole RGB cube: the choice changes about 2 bytes in 100,000 (1102 of 50.3 M over all 2^24 colours).  Where it
    does, a value sat on a rounding boundary of the 8-bit XYB quantiser (xyb.rs:192-199) and one level of X moves the
    decoded colour by up to ~22 code values -- so byte-exactness against the crate is a property of the libm, not of the
    algorithm, and the GPU kernel follows the oracle's platform-independent definition."""
    g = np.arange(256, dtype=np.uint8)
    total = flips = worst = 0
    for r in range(0, 256, 5):
        cube = np.stack(np.meshgrid(np.array([r], np.uint8), g, g, indexing="ij"), -1).reshape(-1, 3)
        a = O.xyb_roundtrip(cube, cube.shape[0], 1)
        b = O.xyb_roundtrip_libm(cube, cube.shape[0], 1)
        d = np.abs(a.astype(int) - b.astype(int))
        flips += int((d > 0).sum())
        total += d.size
        worst = max(worst, int(d.max()))
    assert total == 52 * 65536 * 3
    assert flips / total < 1e-4, (flips, total)
    assert worst <= 40
    # the reference's own three tests (xyb.rs:260-301) hold for either choice
    grey = np.full((4, 4, 3), 128, np.uint8)
    for fn in (O.xyb_roundtrip, O.xyb_roundtrip_libm):
        out = fn(grey, 4, 4).reshape(-1, 3).astype(int)
        assert np.abs(out - 128).max() <= 3
