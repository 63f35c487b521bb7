"""Second, independent restatement of SSIMULACRA2 and DSSIM in numpy ("two independent restatements
... so that a transcription slip in one shows up as a disagreement").  TEST INFRASTRUCTURE ONLY.

Written from the algorithm statements in SURVEY.md Appendix A.3 / A.4 (the published upstream algorithms: libjxl
tools/ssimulacra2.cc == rust-av ssimulacra2 == fast-ssim2; kornelski dssim-core), vectorised over whole planes, not
from oracle/ce_oracle.c.  tests/test_oracle.py::test_two_restatements_agree compares the two on seeded pairs.
PARITY against the real crates stays UNPINNED (they are not vendored): agreement of two
restatements of the same text catches transcription slips, not misreadings of upstream.

fp32 planes with one IEEE rounding per numpy operation; fused multiply-adds of the statement are emulated by forming
the product and sum in float64 and rounding once to float32.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


_LUT = None


def srgb8_to_linear(img_u8: np.ndarray) -> np.ndarray:
    """[h,w,3] uint8 -> [3,h,w] float32 (A.3.1 / src/metrics/dssim.rs:77-85, evaluated in f32 per byte value)."""
    global _LUT
    if _LUT is None:
        import ctypes

        libm = ctypes.CDLL("libm.so.6")   # Rust's f32::powf is the platform libm's powf
        libm.powf.restype = ctypes.c_float
        libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
        lut = np.empty(256, np.float32)
        for v in range(256):
            s = F(v) / F(255.0)
            lut[v] = s / F(12.92) if s <= F(0.04045) else F(libm.powf(float((s + F(0.055)) / F(1.055)), 2.4))
        _LUT = lut
    return np.ascontiguousarray(np.moveaxis(_LUT[img_u8], -1, 0))


# ------------------------------------------------------------------ SSIMULACRA2 (A.3)
_S2_W = [0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0, 0.0004371155730107379, 0.0,
         1.1041726426657346, 0.00066284834129271, 0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0,
         1.8422455520539298, 11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
         1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072, 0.9677937080626833, 0.0,
         0.00014003424285435884, 0.9981766977854967, 0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0,
         0.0013648766163243398, 0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
         0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296, 1.027889937768264,
         225.20515300849274, 0.0, 0.0, 19.213238186143016, 0.0011401524586618361, 0.001237755635509985,
         176.39317598450694, 0.0, 0.0, 24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
         34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0008680556573291698, 0.0, 0.0, 0.0,
         0.0, 0.0, 0.0005313191874358747, 0.0, 0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0004179171803251336,
         0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862, 23.19243343998926, 0.0,
         95.1080498811086, 0.9863978034400682, 0.9834382792465353, 0.0012286405048278493, 171.2667255897307,
         0.9807858872435379, 0.0, 0.0, 0.0, 0.0005130064588990679, 0.0, 0.00010854057858411537]


def _rgauss_coeffs(sigma=1.5):
    """A.3.4 (libjxl CreateRecursiveGaussian), f64 then cast to f32: returns N, mul_in[3], mul_prev[3], mul_prev2[3]."""
    N = int(round(3.2795 * sigma + 0.2546))
    om = [k * math.pi / (2 * N) for k in (1, 3, 5)]
    p = [1 / math.tan(om[0] / 2), -1 / math.tan(om[1] / 2), 1 / math.tan(om[2] / 2)]
    r = [p[0] ** 2 / math.sin(om[0]), -p[1] ** 2 / math.sin(om[1]), p[2] ** 2 / math.sin(om[2])]
    rho = [math.exp(-sigma * sigma * o * o / 2) / N for o in om]
    D13, D35, D51 = p[0] * r[1] - r[0] * p[1], p[1] * r[2] - r[1] * p[2], p[2] * r[0] - r[2] * p[0]
    z15, z35 = D35 / D13, D51 / D13
    A = np.array([[p[0], p[1], p[2]], [r[0], r[1], r[2]], [z15, z35, 1.0]])
    gamma = np.array([1.0, N * N - sigma * sigma, z15 * rho[0] + z35 * rho[1] + rho[2]])
    beta = np.linalg.solve(A, gamma)
    n2 = [-beta[k] * math.cos(om[k] * (N + 1)) for k in range(3)]
    d1 = [-2 * math.cos(om[k]) for k in range(3)]
    return N, np.array(n2, np.float32), np.array([-d for d in d1], np.float32), np.array([-1.0] * 3, np.float32)


_RG = None


def _rgauss_lines(x: np.ndarray) -> np.ndarray:
    """Recursive Gaussian along the LAST axis of x [..., len], zero outside; all lines advance together."""
    global _RG
    if _RG is None:
        _RG = _rgauss_coeffs()
    N, mul_in, mul_prev, mul_prev2 = _RG
    length = x.shape[-1]
    out = np.empty_like(x)
    zero = np.zeros(x.shape[:-1], np.float32)
    prev = [zero.copy() for _ in range(3)]
    prev2 = [zero.copy() for _ in range(3)]
    for n in range(-N + 1, length):
        left = x[..., n - N - 1] if n - N - 1 >= 0 else zero
        right = x[..., n + N - 1] if n + N - 1 < length else zero
        s = left + right
        for k in range(3):
            o = s * mul_in[k]                       # o = sum * mul_in
            o = _fma(mul_prev2[k], prev2[k], o)     # o = fma(mul_prev2, prev2, o)
            o = _fma(mul_prev[k], prev[k], o)       # o = fma(mul_prev, prev, o)
            prev2[k], prev[k] = prev[k], o
        if n >= 0:
            out[..., n] = (prev[0] + prev[1]) + prev[2]
    return out


def _blur(plane: np.ndarray) -> np.ndarray:
    """horizontal then vertical (A.3.4)"""
    h = _rgauss_lines(plane)
    return np.ascontiguousarray(_rgauss_lines(np.ascontiguousarray(h.T)).T)


def _cbrtf(x: np.ndarray) -> np.ndarray:
    return np.cbrt(x.astype(np.float64)).astype(np.float32)


def _xyb_positive(lin: np.ndarray):
    r, g, b = lin
    kb = F(0.0037930732552754493)
    m0 = _fma(F(0.30), r, _fma(F(0.622), g, _fma(F(0.078), b, kb)))
    m1 = _fma(F(0.23), r, _fma(F(0.692), g, _fma(F(0.078), b, kb)))
    m2 = _fma(F(0.24342268924547819), r, _fma(F(0.20476744424496821), g, _fma(F(0.5518098665095536), b, kb)))
    c = [_cbrtf(np.maximum(m, F(0))) - _cbrtf(np.asarray(kb)) for m in (m0, m1, m2)]
    X = F(0.5) * (c[0] - c[1])
    Y = F(0.5) * (c[0] + c[1])
    B = (c[2] - Y) + F(0.55)
    return [_fma(X, F(14.0), F(0.42)), Y + F(0.01), B]


def _down2(lin: np.ndarray) -> np.ndarray:
    _, h, w = lin.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    ys0, ys1 = np.minimum(2 * np.arange(oh), h - 1), np.minimum(2 * np.arange(oh) + 1, h - 1)
    xs0, xs1 = np.minimum(2 * np.arange(ow), w - 1), np.minimum(2 * np.arange(ow) + 1, w - 1)
    s = lin[:, ys0][:, :, xs0] + lin[:, ys0][:, :, xs1]
    s = s + lin[:, ys1][:, :, xs0]
    s = s + lin[:, ys1][:, :, xs1]
    return (s * F(0.25)).astype(np.float32)


def ssimulacra2(ref_u8: np.ndarray, dist_u8: np.ndarray) -> float:
    """[h,w,3] uint8 pair -> score (A.3)."""
    l1, l2 = srgb8_to_linear(ref_u8), srgb8_to_linear(dist_u8)
    avg_ssim, avg_edge = [], []   # per scale: [3][2], [3][4]
    for s in range(6):
        _, h, w = l1.shape
        if w < 8 or h < 8:
            break
        if s > 0:
            l1, l2 = _down2(l1), _down2(l2)
            _, h, w = l1.shape
        p1, p2 = _xyb_positive(l1), _xyb_positive(l2)
        npx = float(h * w)
        sc_ssim, sc_edge = [], []
        for c in range(3):
            i1, i2 = p1[c], p2[c]
            mu1, mu2 = _blur(i1), _blur(i2)
            s11, s22, s12 = _blur(i1 * i1), _blur(i2 * i2), _blur(i1 * i2)
            mdiff = mu1 - mu2
            num_m = _fma(mdiff, -mdiff, F(1.0))
            num_s = _fma(F(2.0), s12 - mu1 * mu2, F(0.0009))
            den_s = ((s11 - mu1 * mu1) + (s22 - mu2 * mu2)) + F(0.0009)
            d = np.maximum(1.0 - ((num_m * num_s) / den_s).astype(np.float64), 0.0)
            sc_ssim.append([d.sum() / npx, ((d ** 4).sum() / npx) ** 0.25])
            d1 = (1.0 + np.abs(i2 - mu2).astype(np.float64)) / (1.0 + np.abs(i1 - mu1).astype(np.float64)) - 1.0
            art, det = np.maximum(d1, 0.0), np.maximum(-d1, 0.0)
            sc_edge.append([art.sum() / npx, ((art ** 4).sum() / npx) ** 0.25, det.sum() / npx, ((det ** 4).sum() / npx) ** 0.25])
        avg_ssim.append(sc_ssim)
        avg_edge.append(sc_edge)
    ssim, i = 0.0, 0
    for c in range(3):
        for sc in range(len(avg_ssim)):
            for n in range(2):
                ssim += _S2_W[i] * abs(avg_ssim[sc][c][n]); i += 1
                ssim += _S2_W[i] * abs(avg_edge[sc][c][n]); i += 1
                ssim += _S2_W[i] * abs(avg_edge[sc][c][n + 2]); i += 1
    ssim *= 0.9562382616834844
    ssim = 2.326765642916932 * ssim - 0.020884521182843837 * ssim * ssim + 6.248496625763138e-05 * ssim * ssim * ssim
    return 100.0 - 10.0 * ssim ** 0.6276336467831387 if ssim > 0 else 100.0


# ------------------------------------------------------------------ DSSIM (A.4)
_DS_K = np.array([[0.095332, 0.118095, 0.095332], [0.118095, 0.146293, 0.118095], [0.095332, 0.118095, 0.095332]], np.float32)
_DS_W = [0.028, 0.197, 0.322, 0.298, 0.155]


def _ds_pass(p: np.ndarray) -> np.ndarray:
    """one 3x3 pass, clamp-replicate edges, nine products summed row-major by rows: (a + b) + c with
    a = (v00 k0 + v01 k1) + v02 k2 etc., no fusion"""
    q = np.pad(p, 1, mode="edge")
    h, w = p.shape
    rows = []
    for dy in range(3):
        r = (q[dy:dy + h, 0:w] * _DS_K[dy, 0] + q[dy:dy + h, 1:w + 1] * _DS_K[dy, 1]) + q[dy:dy + h, 2:w + 2] * _DS_K[dy, 2]
        rows.append(r)
    return ((rows[0] + rows[1]) + rows[2]).astype(np.float32)


def _ds_blur(p: np.ndarray) -> np.ndarray:
    return _ds_pass(_ds_pass(p))


def _cbrt_poly(x):
    y = (F(-0.5) * x + F(1.51)) * x + F(0.2)
    for _ in range(2):
        y3 = (y * y) * y
        y = (y * (y3 + F(2.0) * x)) / (F(2.0) * y3 + x)
    return y


def _to_lab(lin: np.ndarray):
    r, g, b = lin
    D65x, D65z = F(0.9505), F(1.089)
    fx = _fma(b, F(0.1805) / D65x, _fma(g, F(0.3576) / D65x, r * (F(0.4124) / D65x)))
    fy = _fma(b, F(0.0722), _fma(g, F(0.7152), r * F(0.2126)))
    fz = _fma(b, F(0.9505) / D65z, _fma(g, F(0.1192) / D65z, r * (F(0.0193) / D65z)))
    eps, k = F(216.0) / F(24389.0), F(24389.0) / (F(27.0) * F(116.0))

    def f(v):
        with np.errstate(all="ignore"):
            return np.where(v > eps, _cbrt_poly(v) - F(16.0) / F(116.0), k * v).astype(np.float32)

    X, Y, Z = f(fx), f(fy), f(fz)
    return [Y * F(1.05), _fma(F(500.0) / F(220.0), X - Y, F(86.2) / F(220.0)), _fma(F(200.0) / F(220.0), Y - Z, F(107.9) / F(220.0))]


def _ds_down(lin: np.ndarray):
    _, h, w = lin.shape
    hh, hw = h // 2, w // 2
    if hw < 4 or hh < 4:
        return None
    c = lin[:, :2 * hh, :2 * hw]
    return ((((c[:, 0::2, 0::2] + c[:, 0::2, 1::2]) + c[:, 1::2, 0::2]) + c[:, 1::2, 1::2]) * F(0.25)).astype(np.float32)


def dssim(ref_u8: np.ndarray, dist_u8: np.ndarray) -> float:
    """[h,w,3] uint8 pair (alpha = 1) -> DSSIM (A.4)."""
    l1, l2 = srgb8_to_linear(ref_u8), srgb8_to_linear(dist_u8)
    scores = []
    for s in range(5):
        if s > 0:
            d1, d2 = _ds_down(l1), _ds_down(l2)
            if d1 is None:
                break
            l1, l2 = d1, d2
        lab1, lab2 = _to_lab(l1), _to_lab(l2)
        terms = {k: [] for k in ("m11", "m22", "m12", "s1", "s2", "s12")}
        for c in range(3):
            a, b = lab1[c], lab2[c]
            if c > 0:
                a, b = _ds_blur(a), _ds_blur(b)
            mu1, mu2 = _ds_blur(a), _ds_blur(b)
            sq1, sq2, cross = _ds_blur(a * a), _ds_blur(b * b), _ds_blur(a * b)
            m11, m22, m12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
            terms["m11"].append(m11); terms["m22"].append(m22); terms["m12"].append(m12)
            terms["s1"].append(sq1 - m11); terms["s2"].append(sq2 - m22); terms["s12"].append(cross - m12)
        third = F(1.0) / F(3.0)
        avg = {k: ((v[0] + v[1]) + v[2]) * third for k, v in terms.items()}
        c1, c2 = F(0.01) * F(0.01), F(0.03) * F(0.03)
        ssim = (_fma(F(2.0), avg["m12"], c1) * _fma(F(2.0), avg["s12"], c2)) / \
               (((avg["m11"] + avg["m22"]) + c1) * ((avg["s1"] + avg["s2"]) + c2))
        m = ssim.astype(np.float64)
        mean = m.sum() / m.size
        a = max(mean, 0.0) ** (0.5 ** s)
        scores.append(1.0 - np.abs(a - m).sum() / m.size)
    ssim = sum(sc * w for sc, w in zip(scores, _DS_W)) / sum(_DS_W[:len(scores)])
    return 1.0 / max(ssim, 2.220446049250313e-16) - 1.0
