"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances (BASELINE.json north_star): PSNR SSE bit-exact; SSIMULACRA2 0.01 absolute; DSSIM 1e-4 relative;
Butteraugli max / 3-norm 1e-3 relative.  The kernels execute the oracle's fp32 operation sequence, so the
observed differences are far smaller; tests assert the contract tolerance AND report the observed gap.
"""
import os

import numpy as np
import pytest

from codec_eval_b200.synth import G, cheap_distort

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pairs.npz")

S2_TOL = 0.01
DS_RTOL = 1e-4
BA_RTOL = 1e-3


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cases(gold):
    for k, row in enumerate(gold["table"]):
        yield k, int(row[1]), int(row[2]), gold[f"ref{k}"], gold[f"dist{k}"], row


# ------------------------------------------------------------------ PSNR
def test_psnr_sse_bit_exact_golden(gpu, O, gold):
    for k, w, h, ref, dist, row in _cases(gold):
        assert gpu.calculate_sse(ref, dist, w, h) == int(row[5]) == O.sse(ref, dist)
        assert gpu.calculate_psnr(ref, dist, w, h) == row[6]


def test_psnr_reference_rows(gpu):
    data = np.full(100 * 100 * 3, 128, np.uint8)
    assert np.isinf(gpu.calculate_psnr(data, data, 100, 100))
    a = np.full(100 * 100 * 3, 100, np.uint8)
    b = np.full(100 * 100 * 3, 110, np.uint8)
    assert 28.0 < gpu.calculate_psnr(a, b, 100, 100) < 29.0
    with pytest.raises(AssertionError):
        gpu.calculate_psnr(a, b[:-3], 100, 100)


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (17, 9), (333, 77), (1024, 1024), (3840, 2160)])
def test_psnr_sse_random_sizes(gpu, O, w, h):
    rng = np.random.default_rng(w * 7 + h)
    a = rng.integers(0, 256, w * h * 3, dtype=np.uint8)
    b = rng.integers(0, 256, w * h * 3, dtype=np.uint8)
    assert gpu.calculate_sse(a, b, w, h) == O.sse(a, b)
    # extremes: max difference everywhere
    z = np.zeros(w * h * 3, np.uint8)
    f = np.full(w * h * 3, 255, np.uint8)
    assert gpu.calculate_sse(z, f, w, h) == w * h * 3 * 65025


# ------------------------------------------------------------------ XYB round trip
def test_xyb_roundtrip_exact(gpu, O, gold):
    g = np.arange(0, 256, 16, dtype=np.uint8)
    cube = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    out = gpu.xyb_roundtrip(cube, cube.shape[0], 1)
    exp = O.xyb_roundtrip(cube, cube.shape[0], 1)
    assert np.array_equal(out, exp)
    md = np.abs(out.reshape(-1, 3).astype(int) - cube.astype(int)).max()
    assert 0 < md <= 30
    rng = np.random.default_rng(5)
    rnd = rng.integers(0, 256, 200 * 100 * 3, dtype=np.uint8)
    out = gpu.xyb_roundtrip(rnd, 200, 100)
    exp = O.xyb_roundtrip(rnd, 200, 100)
    nbad = int((out != exp).sum())
    assert nbad == 0, f"{nbad} of {rnd.size} bytes differ, max {np.abs(out.astype(int) - exp.astype(int)).max()}"
    assert np.array_equal(gpu.xyb_roundtrip(gold["ref0"], 64, 64).reshape(64, 64, 3), gold["xyb_rt0"])


# ------------------------------------------------------------------ sRGB -> linear
def test_rgb8_to_dssim_image(gpu, O):
    img = gpu.rgb8_to_dssim_image(np.array([255, 255, 255, 0, 0, 0], np.uint8), 2, 1)
    assert abs(img[0, 0, 0] - 1.0) < 1e-3 and img[0, 0, 3] == 1.0 and abs(img[0, 1, 0]) < 1e-3
    img = gpu.rgba8_to_dssim_image(np.array([255, 255, 255, 255, 0, 0, 0, 128], np.uint8), 2, 1)
    assert abs(img[0, 0, 3] - 1.0) < 1e-3 and abs(img[0, 1, 3] - 0.502) < 0.01
    ramp = np.repeat(np.arange(256, dtype=np.uint8), 3)
    assert np.array_equal(gpu.rgb8_to_dssim_image(ramp, 256, 1), O.rgb8_to_dssim_image(ramp, 256, 1))


# ------------------------------------------------------------------ SSIMULACRA2
def test_ssim2_stage_planes(gpu, O, gold):
    """scale-0 planes: positive XYB, mu1, mu2, s11, s22, s12 -- localises a mismatch to the colour kernel,
    the row pass or the column pass."""
    import ctypes as C

    for k in (0, 2):
        w, h = int(gold["table"][k][1]), int(gold["table"][k][2])
        ref, dist = np.ascontiguousarray(gold[f"ref{k}"]), np.ascontiguousarray(gold[f"dist{k}"])
        exp = O.ssimulacra2_scale0_planes(ref, dist, w, h)
        got = np.empty((3, 7, h, w), np.float32)
        st = gpu._L.ce_debug_ssim2_scale0_planes(gpu._h, ref.ctypes.data, dist.ctypes.data, w, h, got.ctypes.data)
        assert st == 0, gpu.last_error()
        names = ["i1", "i2", "mu1", "mu2", "s11", "s22", "s12"]
        for c in range(3):
            for p in range(7):
                d = np.abs(got[c, p] - exp[c, p]).max()
                assert d <= 1e-6, f"case {k} channel {c} plane {names[p]} max abs diff {d}"


def test_ssim2_scale_sums(gpu, O, gold):
    import ctypes as C

    for k, w, h, ref, dist, row in _cases(gold):
        ref, dist = np.ascontiguousarray(ref), np.ascontiguousarray(dist)
        _, exp = O.ssimulacra2_ex(ref, dist, w, h)
        got = np.zeros(6 * 18, np.float64)
        ns = C.c_int()
        st = gpu._L.ce_debug_ssim2_sums(gpu._h, ref.ctypes.data, dist.ctypes.data, w, h,
                                        got.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ns))
        assert st == 0, gpu.last_error()
        assert ns.value == exp.shape[0]
        got = got.reshape(6, 18)[: ns.value]
        err = np.abs(got - exp) / np.maximum(np.abs(exp), 1e-12)
        # The device cube root is the correctly rounded one for all but ~5 values per million (ce_common.cuh cbrt_pos),
        # where it is one ulp off; a single such value moves a pooled fourth-power sum by up to ~1e-4 relative (case 1:
        # 4.5e-5 on the B-channel d^4 sum of scale 1, reproduced on the CPU with the same operation sequence).  Everything
        # else agrees to 1e-6, so the median error stays there.
        assert err.max() < 1e-3, f"case {k}: worst scale/feature {np.unravel_index(err.argmax(), err.shape)} rel {err.max()}"
        assert np.median(err) < 1e-6, f"case {k}: median rel {np.median(err)}"


def test_ssim2_scores_golden(gpu, O, gold):
    worst = 0.0
    for k, w, h, ref, dist, row in _cases(gold):
        got = gpu.calculate_ssimulacra2(ref, dist, w, h)
        worst = max(worst, abs(got - row[7]))
        assert abs(got - row[7]) < S2_TOL, (k, got, row[7])
    print("ssimulacra2 worst |gpu-oracle| =", worst)
    assert worst < 1e-4


def test_ssim2_reference_rows(gpu):
    ref = (np.arange(100 * 100 * 3) % 256).astype(np.uint8)
    assert gpu.calculate_ssimulacra2(ref, ref, 100, 100) > 99.0
    a = np.full(100 * 100 * 3, 100, np.uint8)
    b = np.full(100 * 100 * 3, 200, np.uint8)
    assert gpu.calculate_ssimulacra2(a, b, 100, 100) < 80.0
    from codec_eval_b200.metrics import DimensionMismatch, MetricCalculation

    with pytest.raises(DimensionMismatch):   # ssimulacra2.rs:177-182
        gpu.calculate_ssimulacra2(np.zeros(50 * 50 * 3, np.uint8), np.zeros(100 * 100 * 3, np.uint8), 50, 50)
    with pytest.raises(MetricCalculation):   # len != w*h*3
        gpu.calculate_ssimulacra2(np.zeros(50 * 50 * 3, np.uint8), np.zeros(50 * 50 * 3, np.uint8), 60, 50)
    with pytest.raises(MetricCalculation):   # below 8x8
        gpu.calculate_ssimulacra2(np.zeros(7 * 9 * 3, np.uint8), np.zeros(7 * 9 * 3, np.uint8), 7, 9)


@pytest.mark.parametrize("w,h", [(8, 8), (9, 33), (100, 100), (513, 255), (768, 512)])
def test_ssim2_shapes_vs_oracle(gpu, O, w, h):
    ref = G(w + h, w, h)
    dist = cheap_distort(ref, 70, seed=w)
    got, exp = gpu.calculate_ssimulacra2(ref, dist, w, h), O.ssimulacra2(ref, dist, w, h)
    assert abs(got - exp) < S2_TOL, (got, exp)
    assert gpu.calculate_ssimulacra2(ref, ref, w, h) == 100.0


# ------------------------------------------------------------------ DSSIM
def test_dssim_scales_and_map(gpu, O, gold):
    import ctypes as C

    for k in (0, 2, 3):
        w, h = int(gold["table"][k][1]), int(gold["table"][k][2])
        ref, dist = np.ascontiguousarray(gold[f"ref{k}"]), np.ascontiguousarray(gold[f"dist{k}"])
        exp, exp_sc, exp_map = O.dssim_ex(ref, dist, w, h)
        sc = np.zeros(5, np.float64)
        ns = C.c_int()
        m = np.empty((h, w), np.float32)
        st = gpu._L.ce_debug_dssim_scales(gpu._h, ref.ctypes.data, dist.ctypes.data, w, h,
                                          sc.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ns), m.ctypes.data)
        assert st == 0, gpu.last_error()
        assert ns.value == len(exp_sc)
        assert np.abs(m - exp_map).max() <= 1e-6, f"case {k}: ssim map max abs diff {np.abs(m - exp_map).max()}"
        assert np.abs(sc[: ns.value] - exp_sc).max() < 1e-9, (sc, exp_sc)


def test_dssim_golden(gpu, gold):
    worst = 0.0
    for k, w, h, ref, dist, row in _cases(gold):
        got = gpu.calculate_dssim_rgb8(ref, dist, w, h)
        worst = max(worst, rel(got, row[8]))
        assert rel(got, row[8]) < DS_RTOL, (k, got, row[8])
    print("dssim worst rel =", worst)
    assert worst < 1e-6


def test_dssim_reference_rows(gpu, O):
    from codec_eval_b200.metrics import DimensionMismatch

    a = np.full((100, 100, 4), 0.5, np.float32)
    a[..., 3] = 1.0
    assert gpu.calculate_dssim(a, a) < 1e-4
    b = np.full((100, 100, 4), 0.3, np.float32)
    c = np.full((100, 100, 4), 0.7, np.float32)
    b[..., 3] = c[..., 3] = 1.0
    got = gpu.calculate_dssim(b, c)
    assert got > 0.0 and rel(got, O.dssim_rgbaf32(b, c, 100, 100)) < DS_RTOL
    with pytest.raises(DimensionMismatch):  # dssim.rs:226-249
        gpu.calculate_dssim(np.zeros((100, 100, 4), np.float32), np.zeros((50, 50, 4), np.float32))


def test_dssim_alpha_path(gpu, O):
    rng = np.random.default_rng(3)
    a = rng.random((40, 56, 4), dtype=np.float32)
    b = np.clip(a + rng.normal(0, 0.03, a.shape).astype(np.float32), 0, 1).astype(np.float32)
    got, exp = gpu.calculate_dssim(a, b), O.dssim_rgbaf32(a, b, 56, 40)
    assert rel(got, exp) < DS_RTOL, (got, exp)


@pytest.mark.parametrize("w,h", [(8, 8), (9, 33), (100, 100), (513, 255), (768, 512)])
def test_dssim_shapes_vs_oracle(gpu, O, w, h):
    ref = G(w + h, w, h)
    dist = cheap_distort(ref, 70, seed=w)
    got, exp = gpu.calculate_dssim_rgb8(ref, dist, w, h), O.dssim(ref, dist, w, h)
    assert rel(got, exp) < DS_RTOL, (got, exp)
    assert gpu.calculate_dssim_rgb8(ref, ref, w, h) == 0.0


# ------------------------------------------------------------------ Butteraugli
@pytest.mark.parametrize("sigma", [1.2, 1.56416327805, 2.7, 3.22489901262, 7.15593339443])
def test_ba_blur_stage(gpu, O, sigma):
    rng = np.random.default_rng(11)
    for (w, h) in [(40, 24), (200, 130), (13, 70)]:
        p = (rng.random((h, w), dtype=np.float32) * 100).astype(np.float32)
        out = np.empty_like(p)
        st = gpu._L.ce_debug_ba_blur(gpu._h, p.ctypes.data, w, h, sigma, out.ctypes.data)
        assert st == 0, gpu.last_error()
        exp = O.ba_blur(p, sigma)
        d = np.abs(out - exp).max()
        assert d <= 1e-5, f"sigma {sigma} {w}x{h}: max abs diff {d} at {np.unravel_index(np.abs(out - exp).argmax(), p.shape)}"


def test_ba_opsin_and_psycho_stages(gpu, O, gold):
    for k in (0, 2):
        w, h = int(gold["table"][k][1]), int(gold["table"][k][2])
        ref = np.ascontiguousarray(gold[f"dist{k}"])
        exp = O.butteraugli_opsin(ref, w, h)
        got = np.empty((3, h, w), np.float32)
        st = gpu._L.ce_debug_butteraugli_opsin(gpu._h, ref.ctypes.data, w, h, 80.0, got.ctypes.data)
        assert st == 0, gpu.last_error()
        for c in range(3):
            d = np.abs(got[c] - exp[c]).max()
            assert d <= 1e-4 * max(1.0, np.abs(exp[c]).max()), f"opsin plane {c}: {d}"
        exp = O.butteraugli_psycho(ref, w, h)
        got = np.empty((10, h, w), np.float32)
        st = gpu._L.ce_debug_butteraugli_psycho(gpu._h, ref.ctypes.data, w, h, 80.0, got.ctypes.data)
        assert st == 0, gpu.last_error()
        names = ["lf_x", "lf_y", "lf_b", "mf_x", "mf_y", "mf_b", "hf_x", "hf_y", "uhf_x", "uhf_y"]
        for p in range(10):
            d = np.abs(got[p] - exp[p]).max()
            assert d <= 1e-4 * max(1.0, np.abs(exp[p]).max()), f"case {k} psycho plane {names[p]}: max abs diff {d}"


def test_ba_diffmap(gpu, O, gold):
    for k in (0, 2, 4):
        w, h = int(gold["table"][k][1]), int(gold["table"][k][2])
        ref, dist = np.ascontiguousarray(gold[f"ref{k}"]), np.ascontiguousarray(gold[f"dist{k}"])
        _, _, exp = O.butteraugli_ex(ref, dist, w, h)
        got = np.empty((h, w), np.float32)
        st = gpu._L.ce_debug_butteraugli_diffmap(gpu._h, ref.ctypes.data, dist.ctypes.data, w, h, 80.0, got.ctypes.data)
        assert st == 0, gpu.last_error()
        d = np.abs(got - exp)
        assert d.max() <= 1e-4 * exp.max(), f"case {k}: diffmap max abs diff {d.max()} at {np.unravel_index(d.argmax(), d.shape)} (max value {exp.max()})"


def test_butteraugli_golden(gpu, gold):
    worst = 0.0
    for k, w, h, ref, dist, row in _cases(gold):
        mx, pn = gpu.calculate_butteraugli(ref, dist, w, h, return_pnorm=True)
        worst = max(worst, rel(mx, row[9]), rel(pn, row[10]))
        assert rel(mx, row[9]) < BA_RTOL and rel(pn, row[10]) < BA_RTOL, (k, mx, row[9], pn, row[10])
    print("butteraugli worst rel =", worst)
    assert worst < 1e-5


def test_butteraugli_reference_rows(gpu, O):
    from codec_eval_b200.metrics import DimensionMismatch, MetricCalculation

    ref = (np.arange(100 * 100 * 3) % 256).astype(np.uint8)
    assert gpu.calculate_butteraugli(ref, ref, 100, 100) < 0.01
    a = np.full(100 * 100 * 3, 100, np.uint8)
    b = np.full(100 * 100 * 3, 200, np.uint8)
    assert gpu.calculate_butteraugli(a, b, 100, 100) > 1.0
    assert gpu.calculate_butteraugli_with_intensity(ref, ref, 100, 100, 250.0) < 0.01
    with pytest.raises(DimensionMismatch):
        gpu.calculate_butteraugli(np.zeros(50 * 50 * 3, np.uint8), np.zeros(100 * 100 * 3, np.uint8), 50, 50)
    with pytest.raises(MetricCalculation):
        gpu.calculate_butteraugli(np.zeros(4 * 4 * 3, np.uint8), np.zeros(4 * 4 * 3, np.uint8), 4, 4)
    d = cheap_distort(G(1, 100, 100), 60)
    got = gpu.calculate_butteraugli_with_intensity(G(1, 100, 100), d, 100, 100, 250.0)
    assert rel(got, O.butteraugli(G(1, 100, 100), d, 100, 100, 250.0)[0]) < BA_RTOL


@pytest.mark.parametrize("w,h", [(8, 8), (15, 17), (100, 100), (513, 255), (768, 512)])
def test_butteraugli_shapes_vs_oracle(gpu, O, w, h):
    ref = G(w + h, w, h)
    dist = cheap_distort(ref, 70, seed=w)
    mx, pn = gpu.calculate_butteraugli(ref, dist, w, h, return_pnorm=True)
    emx, epn = O.butteraugli(ref, dist, w, h)
    assert rel(mx, emx) < BA_RTOL and rel(pn, epn) < BA_RTOL, (mx, emx, pn, epn)
    assert gpu.calculate_butteraugli(ref, ref, w, h) == 0.0


# ------------------------------------------------------------------ batched entry
def test_evaluate_batch_mixed(gpu, O, gold):
    from codec_eval_b200.metrics import MetricConfig

    pairs = []
    for k, w, h, ref, dist, row in _cases(gold):
        pairs.append((ref, dist, w, h))
    pairs.append((gold["ref0"], gold["ref0"], 64, 64))            # identical
    pairs.append((gold["ref1"], gold["dist1"], 96, 80))           # duplicate size group
    res = gpu.evaluate_batch(pairs, MetricConfig.all())
    for k, row in enumerate(gold["table"]):
        r = res[k]
        assert r.sse == int(row[5]) and r.psnr == row[6]
        assert abs(r.ssimulacra2 - row[7]) < S2_TOL
        assert rel(r.dssim, row[8]) < DS_RTOL
        assert rel(r.butteraugli, row[9]) < BA_RTOL and rel(r.butteraugli_pnorm3, row[10]) < BA_RTOL
    ident = res[len(gold["table"])]
    assert ident.sse == 0 and np.isinf(ident.psnr) and ident.ssimulacra2 == 100.0 and ident.dssim == 0.0 and ident.butteraugli == 0.0
    assert res[-1].ssimulacra2 == res[1].ssimulacra2 and res[-1].dssim == res[1].dssim


def test_evaluate_batch_status_isolated(gpu, gold):
    from codec_eval_b200 import _lib
    from codec_eval_b200.metrics import MetricConfig

    good = (gold["ref0"], gold["dist0"], 64, 64)
    bad_len = (gold["ref0"], gold["dist1"], 64, 64)
    bad_size = (gold["ref0"], gold["dist0"], 65, 64)
    out = gpu.evaluate_batch_raw([good, bad_len, bad_size, good], MetricConfig.all())
    assert [out[i].status for i in range(4)] == [0, _lib.CE_ERR_DIMENSION_MISMATCH, _lib.CE_ERR_METRIC_CALCULATION, 0]
    assert out[0].ssimulacra2 == out[3].ssimulacra2 and out[0].valid == 15 and out[1].valid == 0
    assert len(gpu.evaluate_batch([], MetricConfig.all())) == 0


def test_xyb_roundtrip_config_applies_to_reference_only(gpu, O, gold):
    from codec_eval_b200.metrics import MetricConfig

    ref, dist = gold["ref0"], gold["dist0"]
    r = gpu.evaluate_batch([(ref, dist, 64, 64)], MetricConfig.all().with_xyb_roundtrip())[0]
    rt = O.xyb_roundtrip(ref, 64, 64)
    assert r.sse == O.sse(rt, dist)
    assert abs(r.ssimulacra2 - O.ssimulacra2(rt, dist, 64, 64)) < S2_TOL
    assert rel(r.dssim, O.dssim(rt, dist, 64, 64)) < DS_RTOL
    assert rel(r.butteraugli, O.butteraugli(rt, dist, 64, 64)[0]) < BA_RTOL


def test_reference_handle(gpu, gold):
    from codec_eval_b200.metrics import GpuReference, MetricConfig

    ref = GpuReference(gpu, gold["ref3"], 256, 256, MetricConfig.ssimulacra2_only())
    a = ref.compare(gold["dist3"]).ssimulacra2
    many = ref.compare_many([gold["dist3"], gold["ref3"]])
    assert a == many[0].ssimulacra2 == gpu.calculate_ssimulacra2(gold["ref3"], gold["dist3"], 256, 256)
    assert many[1].ssimulacra2 == 100.0
    ref.close()


def test_device_batch_sub_batching_and_determinism(gpu, O):
    """A device-resident uniform batch larger than one sub-batch of a small workspace; two runs bit-identical."""
    import torch

    from codec_eval_b200.metrics import GpuMetrics, MetricConfig

    w, h, n = 128, 96, 24
    refs = np.stack([G(i, w, h) for i in range(n)])
    dists = np.stack([cheap_distort(refs[i], 40 + 2 * i, seed=i) for i in range(n)])
    d_ref = torch.from_numpy(refs).cuda()
    d_dist = torch.from_numpy(dists).cuda()
    small = GpuMetrics(0, workspace_bytes=64 << 20)   # forces several sub-batches for Butteraugli
    try:
        small.set_stream(torch.cuda.current_stream().cuda_stream)
        a = small.evaluate_batch_device(d_ref.data_ptr(), d_dist.data_ptr(), n, w, h, MetricConfig.all())
        b = small.evaluate_batch_device(d_ref.data_ptr(), d_dist.data_ptr(), n, w, h, MetricConfig.all())
        assert small.launch_count() > 0
    finally:
        small.close()
    for i in range(n):
        assert (a[i].sse, a[i].ssimulacra2, a[i].dssim, a[i].butteraugli, a[i].butteraugli_pnorm3) == \
               (b[i].sse, b[i].ssimulacra2, b[i].dssim, b[i].butteraugli, b[i].butteraugli_pnorm3)
    for i in (0, 7, 23):
        assert a[i].sse == O.sse(refs[i], dists[i])
        assert abs(a[i].ssimulacra2 - O.ssimulacra2(refs[i], dists[i], w, h)) < S2_TOL
        assert rel(a[i].dssim, O.dssim(refs[i], dists[i], w, h)) < DS_RTOL
        assert rel(a[i].butteraugli, O.butteraugli(refs[i], dists[i], w, h)[0]) < BA_RTOL


def test_shared_reference_batches_match_independent_pairs(gpu, O):
    """evaluate_image's shape: one reference against several distortions.  Pairs that share the reference buffer are
    grouped (reference-side work once); the scores must be bit-identical to evaluating every pair on its own."""
    import torch

    from codec_eval_b200.metrics import MetricConfig

    w, h = 160, 96
    refs = [G(i, w, h) for i in range(3)]
    pairs, ref_index, dists = [], [], []
    for i, r in enumerate(refs):
        for q in (35, 60, 85):
            d = cheap_distort(r, q, seed=10 * i + q)
            pairs.append((r, d, w, h))
            ref_index.append(i)
            dists.append(d)
    cfg = MetricConfig.all()
    grouped = gpu.evaluate_batch(pairs, cfg)
    single = [gpu.evaluate_batch([p], cfg)[0] for p in pairs]
    for a, b in zip(grouped, single):
        assert (a.sse, a.psnr, a.ssimulacra2, a.dssim, a.butteraugli, a.butteraugli_pnorm3) == \
               (b.sse, b.psnr, b.ssimulacra2, b.dssim, b.butteraugli, b.butteraugli_pnorm3)
    # same through the device-resident grouped entry, references in a shuffled order
    order = [2, 0, 1]
    d_ref = torch.from_numpy(np.stack([refs[k] for k in order])).cuda()
    d_dist = torch.from_numpy(np.stack(dists)).cuda()
    ri = [order.index(i) for i in ref_index]
    out = gpu.evaluate_batch_device_grouped(d_ref.data_ptr(), 3, d_dist.data_ptr(), len(dists), ri, w, h, cfg)
    torch.cuda.synchronize()
    for k, b in enumerate(single):
        assert out[k].status == 0 and out[k].sse == b.sse and out[k].ssimulacra2 == b.ssimulacra2
        assert out[k].dssim == b.dssim and out[k].butteraugli == b.butteraugli
    # with the XYB round-trip applied to the (distinct) references
    g2 = gpu.evaluate_batch(pairs, cfg.with_xyb_roundtrip())
    s2 = gpu.evaluate_batch([pairs[4]], cfg.with_xyb_roundtrip())[0]
    assert (g2[4].sse, g2[4].ssimulacra2, g2[4].dssim, g2[4].butteraugli) == (s2.sse, s2.ssimulacra2, s2.dssim, s2.butteraugli)
    rt = O.xyb_roundtrip(refs[1], w, h)
    assert g2[4].sse == O.sse(rt, dists[4])
    # out-of-range index is rejected
    from codec_eval_b200.metrics import CudaError

    with pytest.raises((AssertionError, CudaError)):
        gpu.evaluate_batch_device_grouped(d_ref.data_ptr(), 3, d_dist.data_ptr(), 2, [0, 3], w, h, cfg)


# ------------------------------------------------------------------ BASELINE.json full sizes
def test_full_size_1024_all_metrics_vs_oracle(gpu, O):
    """cfg5 shape (1024x1024, all metrics): one pair against the oracle, plus batch invariance."""
    from codec_eval_b200.metrics import MetricConfig

    w = h = 1024
    ref = G(5, w, h)
    dist = cheap_distort(ref, 75, seed=5)
    r = gpu.evaluate_batch([(ref, dist, w, h)], MetricConfig.all())[0]
    assert r.sse == O.sse(ref, dist)
    assert abs(r.ssimulacra2 - O.ssimulacra2(ref, dist, w, h)) < S2_TOL
    assert rel(r.dssim, O.dssim(ref, dist, w, h)) < DS_RTOL
    emx, epn = O.butteraugli(ref, dist, w, h)
    assert rel(r.butteraugli, emx) < BA_RTOL and rel(r.butteraugli_pnorm3, epn) < BA_RTOL
    # the same pair inside a larger shared-reference batch gives the same bits
    d2 = cheap_distort(ref, 50, seed=6)
    many = gpu.evaluate_batch([(ref, d2, w, h), (ref, dist, w, h), (ref, ref, w, h)], MetricConfig.all())
    assert (many[1].sse, many[1].ssimulacra2, many[1].dssim, many[1].butteraugli) == (r.sse, r.ssimulacra2, r.dssim, r.butteraugli)
    assert many[2].sse == 0 and many[2].ssimulacra2 == 100.0 and many[2].dssim == 0.0 and many[2].butteraugli == 0.0
    # stronger distortion scores worse on every metric
    assert many[0].sse > r.sse and many[0].ssimulacra2 < r.ssimulacra2 and many[0].dssim > r.dssim and many[0].butteraugli > r.butteraugli


def test_full_size_4k_dssim_butteraugli_vs_oracle(gpu, O):
    """cfg4 shape (3840x2160, Butteraugli + DSSIM): odd pyramid levels (2160 -> ... -> 135 -> 67), one pair against the
    oracle; identical pair is exactly zero."""
    from codec_eval_b200.metrics import MetricConfig

    w, h = 3840, 2160
    ref = G(9, w, h)
    dist = cheap_distort(ref, 85, seed=9)
    cfg = MetricConfig(dssim=True, butteraugli=True, psnr=True)
    r = gpu.evaluate_batch([(ref, dist, w, h), (ref, ref, w, h)], cfg)
    assert r[0].sse == O.sse(ref, dist)
    assert rel(r[0].dssim, O.dssim(ref, dist, w, h)) < DS_RTOL
    emx, epn = O.butteraugli(ref, dist, w, h)
    assert rel(r[0].butteraugli, emx) < BA_RTOL and rel(r[0].butteraugli_pnorm3, epn) < BA_RTOL
    assert r[0].ssimulacra2 is None
    assert r[1].sse == 0 and r[1].dssim == 0.0 and r[1].butteraugli == 0.0 and np.isinf(r[1].psnr)


def test_codec_iter_sweep_shape_ssim2_only(gpu, O):
    """cfg3 shape: 512x512 references x 3 qualities x {plain, XYB round-tripped reference}, SSIMULACRA2 only, through
    the reference handle (Ssimulacra2Reference::new / .compare, crates/codec-iter/src/eval.rs:138-149)."""
    from codec_eval_b200.metrics import GpuReference, MetricConfig

    w = h = 512
    for i in range(2):
        ref = G(20 + i, w, h)
        dists = [cheap_distort(ref, q, seed=i) for q in (75, 85, 95)]
        for cfg in (MetricConfig.ssimulacra2_only(), MetricConfig.ssimulacra2_only().with_xyb_roundtrip()):
            handle = GpuReference(gpu, ref, w, h, cfg)
            got = [m.ssimulacra2 for m in handle.compare_many(dists)]
            handle.close()
            base = O.xyb_roundtrip(ref, w, h).reshape(h, w, 3) if cfg.xyb_roundtrip else ref
            for g_, d in zip(got, dists):
                assert abs(g_ - O.ssimulacra2(np.ascontiguousarray(base), d, w, h)) < S2_TOL
            assert got[0] < got[1] < got[2]


def test_many_tiny_pairs_and_workspace_limits(gpu, O):
    """More pairs than one sub-batch may hold (grid.z limits), smallest legal size, and a workspace that cannot hold
    one pair (CE_ERR_OUT_OF_MEMORY, no crash)."""
    import torch

    from codec_eval_b200 import _lib
    from codec_eval_b200.metrics import CudaError, GpuMetrics, MetricConfig

    w = h = 8
    n = 12000
    rng = np.random.default_rng(1)
    refs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    dists = np.clip(refs.astype(int) + rng.integers(-9, 10, refs.shape), 0, 255).astype(np.uint8)
    d_ref, d_dist = torch.from_numpy(refs).cuda(), torch.from_numpy(dists).cuda()
    out = gpu.evaluate_batch_device(d_ref.data_ptr(), d_dist.data_ptr(), n, w, h, MetricConfig.all())
    torch.cuda.synchronize()
    for i in (0, 4999, 9999, 10000, 11999):     # both sides of the sub-batch boundary
        assert out[i].status == 0 and out[i].valid == 15
        assert out[i].sse == O.sse(refs[i], dists[i])
        assert abs(out[i].ssimulacra2 - O.ssimulacra2(refs[i], dists[i], w, h)) < S2_TOL
        assert rel(out[i].dssim, O.dssim(refs[i], dists[i], w, h)) < DS_RTOL
        assert rel(out[i].butteraugli, O.butteraugli(refs[i], dists[i], w, h)[0]) < BA_RTOL
    tiny = GpuMetrics(0, workspace_bytes=8 << 20)
    try:
        big = np.zeros(1024 * 1024 * 3, np.uint8)
        with pytest.raises(CudaError) as e:
            tiny.evaluate_batch([(big, big, 1024, 1024)], MetricConfig.all())
        assert f"status {_lib.CE_ERR_OUT_OF_MEMORY}" in str(e.value)
        # the context is still usable afterwards
        assert tiny.calculate_psnr(big[:192], big[:192], 8, 8) == float("inf")
    finally:
        tiny.close()


def test_repeated_runs_are_bit_identical_with_all_metrics_sharing_the_sms(gpu):
    """Race detector for the asynchronously staged tiles (TMA / cp.async refills of shared-memory buffers): the three
    perceptual metrics run on separate streams, so their kernels share SMs and the load/store queues are congested.
    A bulk copy that overtakes a pending shared-memory load showed up as a Butteraugli score that changed in ~1 of 3
    runs at this size before the proxy fences were added."""
    from codec_eval_b200.metrics import GpuMetrics, MetricConfig

    w, h = 3840, 2160
    ref = G(9, w, h)
    pairs = [(ref, cheap_distort(ref, 50 + 7 * i, seed=i), w, h) for i in range(2)]
    cfg = MetricConfig.all()
    serial = gpu.evaluate_batch_raw(pairs, cfg)        # 16.6 MPix-pairs: above the fork threshold, metrics back to back
    want = [(o.status, o.sse, o.dssim, o.ssimulacra2, o.butteraugli, o.butteraugli_pnorm3) for o in serial[:2]]
    os.environ["CE_FORK"] = "1"                        # read at context creation: always fork
    try:
        forked = GpuMetrics(0, workspace_bytes=12 << 30)
    finally:
        del os.environ["CE_FORK"]
    try:
        for rep in range(20):
            out = forked.evaluate_batch_raw(pairs, cfg)
            cur = [(o.status, o.sse, o.dssim, o.ssimulacra2, o.butteraugli, o.butteraugli_pnorm3) for o in out[:2]]
            assert cur == want, (rep, cur, want)       # and forked == serial, bit for bit
    finally:
        forked.close()


# ------------------------------------------------------------------ JPEG-distorted full sizes (what the bench times)
@pytest.mark.parametrize("w,h,q,ss", [(768, 512, 75, 2), (768, 512, 90, 0), (1024, 1024, 50, 2), (1024, 1024, 95, 2),
                                      (3840, 2160, 85, 2)])
def test_jpeg_distorted_full_sizes_all_metrics_vs_oracle(gpu, O, w, h, q, ss):
    """BASELINE.json's shapes with the distortion the benchmark uses (baseline JPEG, 4:2:0 and 4:4:4, low and high
    quality) rather than the codec-free stand-in: all four metrics, 4K included, at the contract tolerances."""
    from codec_eval_b200.metrics import MetricConfig

    ref = G(w + q, w, h)
    dist = O.jpeg_roundtrip(ref, w, h, q, ss)
    assert np.array_equal(gpu.jpeg_roundtrip(ref, w, h, q, ss), dist)      # the on-device source makes the same image
    r = gpu.evaluate_batch([(ref, dist, w, h)], MetricConfig.all())[0]
    e = O.evaluate_batch(ref[None], dist[None], w, h, 15, threads=0)[0]
    assert r.sse == e.sse and r.psnr == e.psnr
    assert abs(r.ssimulacra2 - e.ssimulacra2) < S2_TOL, (r.ssimulacra2, e.ssimulacra2)
    assert rel(r.dssim, e.dssim) < DS_RTOL, (r.dssim, e.dssim)
    assert rel(r.butteraugli, e.butteraugli) < BA_RTOL and rel(r.butteraugli_pnorm3, e.butteraugli_pnorm3) < BA_RTOL
    print(f"{w}x{h} q{q}: d_ssim2 {abs(r.ssimulacra2 - e.ssimulacra2):.2e} rel_dssim {rel(r.dssim, e.dssim):.2e} "
          f"rel_ba {rel(r.butteraugli, e.butteraugli):.2e}")


def test_ssim2_xyb_planes_follow_the_reference_cube_root(gpu, O):
    """The XYB planes of the GPU path against the oracle's (whose cbrtf restates yuvxyb-math's, i.e. is correctly
    rounded): the device's fp32-only sequence may differ by one ulp in ~5 values per million, never more."""
    w, h = 512, 512
    ref = G(77, w, h)
    dist = O.jpeg_roundtrip(ref, w, h, 90, 0)
    got = np.empty((3, 7, h, w), np.float32)
    import ctypes as C

    st = gpu._L.ce_debug_ssim2_scale0_planes(gpu._h, ref.ctypes.data, dist.ctypes.data, w, h, got.ctypes.data)
    assert st == 0
    exp = O.ssimulacra2_scale0_planes(ref, dist, w, h)
    for c in range(3):
        for k in (0, 1):     # i1, i2: pointwise colour conversion only
            a, b = got[c, k], np.asarray(exp)[c, k]
            assert np.abs(a - b).max() <= 2.4e-7 * 2, (c, k)         # <= 1 ulp of values < 2
            assert (a != b).mean() < 1e-4, (c, k, (a != b).mean())


# ------------------------------------------------------------------ host-side pipeline of ce_evaluate_batch
def test_reference_sharing_needs_equal_pointer_and_ref_id(gpu):
    """ce_pair.ref_id: pairs share a reference when pointer AND ref_id are equal; a different ref_id on the same buffer
    only switches the sharing off (two uploads) -- the scores are the same bits either way."""
    import ctypes as C

    from codec_eval_b200 import _lib
    from codec_eval_b200.metrics import MetricConfig

    w, h = 160, 96
    ref = G(3, w, h)
    dists = [cheap_distort(ref, q, seed=q) for q in (30, 55, 80, 95)]
    n = len(dists)
    cfg = MetricConfig.all()

    def run(ids):
        tab = (_lib.CePair * n)()
        for i, d in enumerate(dists):
            tab[i] = _lib.CePair(ref.ctypes.data, d.ctypes.data, ref.size, d.size, w, h, ids[i], 0)
        l0 = gpu.launch_count()
        out = gpu.evaluate_pair_table(tab, n, cfg)
        return [(o.status, o.sse, o.dssim, o.ssimulacra2, o.butteraugli, o.butteraugli_pnorm3) for o in out[:n]], gpu.launch_count() - l0

    shared, l_shared = run([7, 7, 7, 7])
    apart, l_apart = run([0, 1, 2, 3])
    mixed, _ = run([0, 0, 5, 5])
    assert shared == apart == mixed
    assert all(s[0] == 0 for s in shared)


def test_pageable_registered_and_chunked_host_batches_agree(gpu, O):
    """The same host batch from pageable arrays, from arrays page-locked with ce_host_register, and through a context
    whose workspace holds only a few pairs (many chunks: launch chunk k, stage chunk k+1, finish chunk k): identical
    results in submission order."""
    from codec_eval_b200.metrics import GpuMetrics, MetricConfig

    w, h = 192, 128
    refs = [G(40 + i, w, h) for i in range(5)]
    pairs = []
    for i, r in enumerate(refs):
        for q in (35, 60, 85):
            pairs.append((r, cheap_distort(r, q, seed=i * 7 + q), w, h))
    cfg = MetricConfig.all()
    key = lambda o: (o.status, o.valid, o.sse, o.dssim, o.ssimulacra2, o.butteraugli, o.butteraugli_pnorm3)
    base = [key(o) for o in gpu.evaluate_batch_raw(pairs, cfg)[:len(pairs)]]
    small = GpuMetrics(0, workspace_bytes=48 << 20)
    try:
        cap = small.sub_batch_capacity(cfg, w, h)
        assert 1 <= cap < len(pairs)                                        # several chunks
        assert [key(o) for o in small.evaluate_batch_raw(pairs, cfg)[:len(pairs)]] == base
        big = np.ascontiguousarray(np.stack([p[1] for p in pairs]))       # one allocation, page-locked in place
        small.host_register(big)
        try:
            reg = [(p[0], big[i], w, h) for i, p in enumerate(pairs)]
            assert [key(o) for o in small.evaluate_batch_raw(reg, cfg)[:len(pairs)]] == base
        finally:
            small.host_unregister(big)
    finally:
        small.close()
    assert base[4][2] == O.sse(pairs[4][0], pairs[4][1])
    assert gpu.sub_batch_capacity(cfg, w, h) > cap


def test_reference_handle_outlives_its_context():
    """ce_reference_destroy after ce_ctx_destroy (Python: `with GpuMetrics() as g: ref = GpuReference(g, ...)` and the
    handle's finaliser later) must not touch the freed context."""
    from codec_eval_b200.metrics import GpuMetrics, GpuReference, MetricConfig

    w, h = 64, 48
    ref = G(1, w, h)
    with GpuMetrics(0, workspace_bytes=256 << 20) as g:
        handle = GpuReference(g, ref, w, h, MetricConfig.ssimulacra2_only())
        assert handle.compare(ref).ssimulacra2 == 100.0
        raw = handle._h
    assert handle._h is None           # closed with its context
    # and the C entry itself tolerates the other order
    g2 = GpuMetrics(0, workspace_bytes=256 << 20)
    h2 = GpuReference(g2, ref, w, h, MetricConfig.ssimulacra2_only())
    keep, h2._h = h2._h, None          # detach from the wrapper so that closing the context does not close it
    g2._refs.clear()
    L = g2._L
    g2.close()
    L.ce_reference_destroy(keep)


def test_failing_pair_is_named_after_the_metric_that_failed(gpu):
    from codec_eval_b200.metrics import MetricCalculation, MetricConfig

    tiny = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(MetricCalculation) as e:
        gpu.evaluate_batch([(tiny, tiny, 4, 4)], MetricConfig(butteraugli=True, psnr=True))
    assert e.value.metric == "Butteraugli"
    with pytest.raises(MetricCalculation) as e:
        gpu.evaluate_batch([(tiny, tiny, 4, 4)], MetricConfig.all())
    assert e.value.metric == "SSIMULACRA2"
