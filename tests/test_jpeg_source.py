"""On-device distortion source: baseline-JPEG sample-domain round trip.

CPU part: the C oracle (oracle/ce_oracle_jpeg.c) against a REAL codec -- Pillow's JPEG save -> load (libjpeg-turbo),
bit for bit -- and against committed golden vectors (tests/golden/jpeg_golden.npz, made by tests/golden/make_jpeg_golden.py
from Pillow, so the pin survives a Pillow upgrade).  GPU part: the CUDA kernels, through the C ABI, against the oracle
(bit-exact: integer work) and the sweep entry against evaluate_batch on the same images."""
import os

import numpy as np
import pytest

from codec_eval_b200.synth import G, J

GOLD = os.path.join(os.path.dirname(__file__), "golden", "jpeg_golden.npz")
SIZES = [(1, 1), (8, 8), (9, 33), (15, 17), (13, 70), (17, 16), (64, 48), (100, 60), (77, 35), (160, 96)]
QUALITIES = [1, 5, 30, 50, 75, 90, 95, 100]


def test_qtables_match_pillow(O):
    """jcparam.c quality scaling: the tables Pillow writes into a file are the oracle's."""
    import io

    from PIL import Image

    img = Image.fromarray(G(0, 16, 16), "RGB")
    zig = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
           62, 63]
    for q in QUALITIES:
        buf = io.BytesIO()
        img.save(buf, format="JPEG", quality=q)
        buf.seek(0)
        qt = Image.open(buf).quantization
        for comp in (0, 1):
            got = O.jpeg_qtable(bool(comp), q)
            file_tbl = np.asarray(qt[comp])
            # Pillow >= 8.3 returns the tables in natural order; older ones in zigzag order
            nat = file_tbl if np.array_equal(file_tbl, got) else None
            if nat is None:
                nat = np.zeros(64, np.int64)
                nat[zig] = file_tbl
            assert np.array_equal(nat, got), (q, comp)


@pytest.mark.parametrize("ss", [0, 2])
def test_oracle_is_bit_exact_with_pillow(O, ss):
    for (w, h) in SIZES:
        img = G(w + h, w, h)
        for q in QUALITIES:
            assert np.array_equal(O.jpeg_roundtrip(img, w, h, q, ss), J(img, q, ss)), (w, h, q, ss)


def test_oracle_against_golden(O):
    g = np.load(GOLD)
    for k, (w, h, q, ss) in enumerate(g["table"]):
        out = O.jpeg_roundtrip(g[f"src{k}"], int(w), int(h), int(q), int(ss))
        assert np.array_equal(out, g[f"out{k}"]), (w, h, q, ss)


def test_oracle_bench_shape_768x512(O):
    img = G(3, 768, 512)
    for q in (50, 85):
        for ss in (0, 2):
            assert np.array_equal(O.jpeg_roundtrip(img, 768, 512, q, ss), J(img, q, ss))


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("ss", [0, 2])
def test_cuda_roundtrip_bit_exact(gpu, O, ss):
    for (w, h) in SIZES + [(768, 512), (513, 255)]:
        img = G(w + h, w, h)
        for q in (1, 30, 75, 95, 100):
            got = gpu.jpeg_roundtrip(img, w, h, q, ss)
            exp = O.jpeg_roundtrip(img, w, h, q, ss)
            assert np.array_equal(got, exp), (w, h, q, ss, int(np.abs(got.astype(int) - exp).max()))


@pytest.mark.gpu
def test_cuda_roundtrip_matches_real_codec(gpu):
    img = G(5, 768, 512)
    for q, ss in ((50, 2), (85, 2), (85, 0)):
        assert np.array_equal(gpu.jpeg_roundtrip(img, 768, 512, q, ss), J(img, q, ss))


@pytest.mark.gpu
def test_cuda_roundtrip_golden(gpu):
    g = np.load(GOLD)
    for k, (w, h, q, ss) in enumerate(g["table"]):
        assert np.array_equal(gpu.jpeg_roundtrip(g[f"src{k}"], int(w), int(h), int(q), int(ss)), g[f"out{k}"])


@pytest.mark.gpu
def test_device_roundtrip_many(gpu, O):
    import torch

    w, h, qs = 96, 80, [40, 75, 90]
    refs = np.stack([G(i, w, h) for i in range(5)])
    d_ref = torch.from_numpy(refs).cuda()
    d_out = torch.empty((5 * len(qs), h, w, 3), dtype=torch.uint8, device="cuda")
    gpu.jpeg_roundtrip_device(d_ref.data_ptr(), 5, w, h, qs, 2, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for r in range(5):
        for k, q in enumerate(qs):
            assert np.array_equal(got[r * len(qs) + k], O.jpeg_roundtrip(refs[r], w, h, q, 2)), (r, q)


@pytest.mark.gpu
def test_sweep_equals_batch_on_the_same_images(gpu, O):
    from codec_eval_b200.metrics import MetricConfig

    w, h, qs = 160, 96, [50, 75, 95]
    refs = [G(10 + i, w, h) for i in range(4)]
    cfg = MetricConfig.all()
    table = gpu.evaluate_jpeg_sweep(refs, w, h, qs, cfg, subsampling=2)
    pairs = [(r, J(r, q, 2), w, h) for r in refs for q in qs]
    flat = gpu.evaluate_batch(pairs, cfg)
    for i, r in enumerate(refs):
        for k, q in enumerate(qs):
            a, b = table[i][k], flat[i * len(qs) + k]
            assert a.sse == b.sse and a.psnr == b.psnr
            assert a.ssimulacra2 == b.ssimulacra2 and a.dssim == b.dssim
            assert a.butteraugli == b.butteraugli and a.butteraugli_pnorm3 == b.butteraugli_pnorm3
    # and against the CPU oracle chain (oracle JPEG -> oracle metric)
    d = O.jpeg_roundtrip(refs[0], w, h, 75, 2)
    assert abs(table[0][1].ssimulacra2 - O.ssimulacra2(refs[0], d, w, h)) < 0.01
    assert table[0][1].sse == O.sse(refs[0], d)


@pytest.mark.gpu
def test_sweep_argument_errors(gpu):
    from codec_eval_b200.metrics import MetricConfig

    r = G(0, 16, 16)
    with pytest.raises(AssertionError):
        gpu.jpeg_roundtrip(r, 16, 16, 0, 2)
    with pytest.raises(AssertionError):
        gpu.jpeg_roundtrip(r, 16, 16, 50, 1)
    assert gpu.evaluate_jpeg_sweep([], 16, 16, [50], MetricConfig.fast()) == []


@pytest.mark.gpu
def test_sweep_with_a_small_workspace_sub_batches_and_matches(gpu):
    """A context whose workspace holds only a few pairs must cut the sweep into chunks / sub-batches and still return
    the same bits; 4:4:4 as well as 4:2:0."""
    from codec_eval_b200.metrics import GpuMetrics, MetricConfig

    w, h, qs = 256, 192, [30, 60, 80, 95]
    refs = [G(70 + i, w, h) for i in range(6)]
    cfg = MetricConfig.all()
    for ss in (2, 0):
        big = gpu.evaluate_jpeg_sweep_raw(refs, w, h, qs, cfg, subsampling=ss)
        with GpuMetrics(0, workspace_bytes=96 << 20) as small:
            out = small.evaluate_jpeg_sweep_raw(refs, w, h, qs, cfg, subsampling=ss)
        for i in range(len(refs) * len(qs)):
            a, b = big[i], out[i]
            assert (a.status, a.sse, a.dssim, a.ssimulacra2, a.butteraugli, a.butteraugli_pnorm3) == \
                   (b.status, b.sse, b.dssim, b.ssimulacra2, b.butteraugli, b.butteraugli_pnorm3), (ss, i)
    # monotone in quality for every reference (PSNR of a JPEG round trip)
    for r in range(len(refs)):
        p = [big[r * len(qs) + k].psnr for k in range(len(qs))]
        assert p == sorted(p), p
