"""Second, independent restatement of SSIMULACRA2, DSSIM and Butteraugli in numpy ("two independent restatements
... so that a transcription slip in one shows up as a disagreement").  TEST INFRASTRUCTURE ONLY.

Written from the algorithm statements in SURVEY.md Appendix A.3 / A.4 / A.5 (the published upstream algorithms: libjxl
tools/ssimulacra2.cc == rust-av ssimulacra2 == fast-ssim2; kornelski dssim-core; libjxl butteraugli.cc), vectorised
over whole planes, not from oracle/ce_oracle.c.  tests/test_oracle.py::test_two_restatements_agree compares the two on seeded pairs.
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


# ------------------------------------------------------------------ Butteraugli (A.5)
_MALTA_HF = [
    [(0, i) for i in range(-4, 5)], [(i, 0) for i in range(-4, 5)],
    [(i, i) for i in range(-3, 4)], [(i, -i) for i in range(-3, 4)],
    [(-4, 1), (-3, 1), (-2, 1), (-1, 0), (0, 0), (1, 0), (2, -1), (3, -1), (4, -1)],
    [(-4, -1), (-3, -1), (-2, -1), (-1, 0), (0, 0), (1, 0), (2, 1), (3, 1), (4, 1)],
    [(-1, -4), (-1, -3), (-1, -2), (0, -1), (0, 0), (0, 1), (1, 2), (1, 3), (1, 4)],
    [(1, -4), (1, -3), (1, -2), (0, -1), (0, 0), (0, 1), (-1, 2), (-1, 3), (-1, 4)],
    [(-3, -2), (-2, -1), (-1, -1), (0, 0), (1, 1), (2, 1), (3, 2)], [(-3, 2), (-2, 1), (-1, 1), (0, 0), (1, -1), (2, -1), (3, -2)],
    [(-2, -3), (-1, -2), (-1, -1), (0, 0), (1, 1), (1, 2), (2, 3)], [(-2, 3), (-1, 2), (-1, 1), (0, 0), (1, -1), (1, -2), (2, -3)],
    [(2, -4), (2, -3), (1, -2), (1, -1), (0, 0), (0, 1), (-1, 2), (-1, 3)],
    [(-2, -4), (-2, -3), (-1, -2), (-1, -1), (0, 0), (0, 1), (1, 2), (1, 3)],
    [(-4, -2), (-3, -2), (-2, -1), (-1, -1), (0, 0), (1, 0), (2, 1), (3, 1)],
    [(-4, 2), (-3, 2), (-2, 1), (-1, 1), (0, 0), (1, 0), (2, -1), (3, -1)]]
_MALTA_LF = [
    [(0, -4), (0, -2), (0, 0), (0, 2), (0, 4)], [(-4, 0), (-2, 0), (0, 0), (2, 0), (4, 0)],
    [(-3, -3), (-2, -2), (0, 0), (2, 2), (3, 3)], [(-3, 3), (-2, 2), (0, 0), (2, -2), (3, -3)],
    [(-4, 1), (-2, 1), (0, 0), (2, -1), (4, -1)], [(-4, -1), (-2, -1), (0, 0), (2, 1), (4, 1)],
    [(-1, -4), (-1, -2), (0, 0), (1, 2), (1, 4)], [(1, -4), (1, -2), (0, 0), (-1, 2), (-1, 4)],
    [(-3, -2), (-2, -1), (0, 0), (2, 1), (3, 2)], [(-3, 2), (-2, 1), (0, 0), (2, -1), (3, -2)],
    [(-2, -3), (-1, -2), (0, 0), (1, 2), (2, 3)], [(-2, 3), (-1, 2), (0, 0), (1, -2), (2, -3)],
    [(2, -4), (1, -2), (0, 0), (-1, 2), (-2, 4)], [(-2, -4), (-1, -2), (0, 0), (1, 2), (2, 4)],
    [(-4, -2), (-2, -1), (0, 0), (2, 1), (4, 2)], [(-4, 2), (-2, 1), (0, 0), (2, -1), (4, -2)]]


def _ba_blur(p: np.ndarray, sigma: float) -> np.ndarray:
    """A.5.0: truncated Gaussian, separable; borders renormalised by the in-range taps; sigma 1.2 = mirror-padded 5 taps."""
    diff = max(1, int(2.25 * abs(sigma)))
    wts = np.array([math.exp(-(i * i) / (2.0 * sigma * sigma)) for i in range(-diff, diff + 1)]).astype(np.float32).astype(np.float64)

    def line(a):   # along the last axis
        n = a.shape[-1]
        a64 = a.astype(np.float64)
        if diff == 2:
            wn = (wts.astype(np.float32) * (F(1.0) / wts.astype(np.float32).sum(dtype=np.float32))).astype(np.float64)
            q = np.pad(a64, [(0, 0)] * (a.ndim - 1) + [(2, 2)], mode="symmetric")
            out = sum(wn[t] * q[..., t:t + n] for t in range(5))
            return out.astype(np.float32)
        q = np.pad(a64, [(0, 0)] * (a.ndim - 1) + [(diff, diff)])
        ones = np.pad(np.ones(n), (diff, diff))
        num = sum(wts[t] * q[..., t:t + n] for t in range(2 * diff + 1))
        den = sum(wts[t] * ones[t:t + n] for t in range(2 * diff + 1))
        return (num / den).astype(np.float32)

    hpass = line(p)
    return np.ascontiguousarray(line(np.ascontiguousarray(hpass.T)).T)


def _opsin_abs(r, g, b):
    o0 = F(0.29956550340058319) * r + F(0.63373087833825936) * g + F(0.077705617820981968) * b + F(1.7557483643287353)
    o1 = F(0.22158691104574774) * r + F(0.69391388044116142) * g + F(0.0987313588422) * b + F(1.7557483643287353)
    o2 = F(0.02) * r + F(0.02) * g + F(0.20480129041026129) * b + F(12.226454707163354)
    return o0, o1, o2


def _gamma(v):
    return (19.245013259874995 * np.log(np.maximum(v, 0).astype(np.float64) + 9.9710635769299145) - 23.16046239805755).astype(np.float32)


def _remove_range(v, w):
    return np.where(v > w, v - w, np.where(v < -w, v + w, F(0))).astype(np.float32)


def _amplify_range(v, w):
    return np.where(v > w, v + w, np.where(v < -w, v - w, v + v)).astype(np.float32)


def _max_clamp(v, m):
    k = F(0.724216145665)
    return np.where(v >= m, (v - m) * k + m, np.where(v < -m, (v + m) * k - m, v)).astype(np.float32)


def _ba_psycho(lin: np.ndarray, intensity: float):
    """linear rgb [3,h,w] (0..1) -> dict of frequency bands (A.5.1-3) and the mask input (A.5.6)."""
    rgb = [c * F(intensity) for c in lin]
    blurred = [_ba_blur(c, 1.2) for c in rgb]
    pre = [np.maximum(o, F(1e-4)) for o in _opsin_abs(*blurred)]
    sens = [np.maximum(_gamma(p) / p, F(1e-4)) for p in pre]
    cur = [o * s for o, s in zip(_opsin_abs(*rgb), sens)]
    cur = [np.maximum(cur[0], F(1.7557483643287353)), np.maximum(cur[1], F(1.7557483643287353)), np.maximum(cur[2], F(12.226454707163354))]
    xyb = [cur[0] - cur[1], cur[0] + cur[1], cur[2]]
    lf = [_ba_blur(c, 7.15593339443) for c in xyb]
    mf = [c - l for c, l in zip(xyb, lf)]
    mfb = [_ba_blur(c, 3.22489901262) for c in mf]
    hf = [mf[0] - mfb[0], mf[1] - mfb[1]]
    mfo = [_remove_range(mfb[0], F(0.29)), _amplify_range(mfb[1], F(0.1)), mfb[2]]
    hf[0] = hf[0] * (F(0.653020556257) + F(46.0) * F(1 - 0.653020556257) / (F(46.0) + hf[1] * hf[1]))
    hfb = [_ba_blur(c, 1.56416327805) for c in hf]
    uhf_x = _remove_range(hf[0] - hfb[0], F(0.04))
    hf_x = _remove_range(hfb[0], F(1.5))
    hh = _max_clamp(hfb[1], F(28.4691806922))
    u = _max_clamp(hf[1] - hh, F(5.19175294647))
    uhf_y = F(2.69313763794) * u
    hf_y = _amplify_range(F(2.155) * hh, F(0.132))
    lfb = lf[2] + F(-0.362267051518) * lf[1]
    lfo = [lf[0] * F(33.832837186260), lf[1] * F(14.458268100570), lfb * F(49.87984651440)]
    m = np.sqrt(((uhf_x + hf_x) * F(2.5)) ** 2 + (uhf_y * F(0.4) + hf_y * F(0.4)) ** 2)
    bb = F(6.19424080439) * F(12.61050594197)
    pm = np.sqrt(F(6.19424080439) * np.abs(m) + bb) - np.sqrt(bb)
    return {"lf": lfo, "mf": mfo, "hf": [hf_x, hf_y], "uhf": [uhf_x, uhf_y], "bl": _ba_blur(pm.astype(np.float32), 2.7)}


def _malta(l0, l1, w0gt1, w0lt1, norm1, lf_patterns):
    length, mulli = 3.75, (0.611612573796 if lf_patterns else 0.39905817637)
    wpre0 = mulli * math.sqrt(0.5 * w0gt1) / (2 * length + 1)
    wpre1 = mulli * math.sqrt(0.33 * w0lt1) / (2 * length + 1)
    n20, n21, n1 = F(wpre0 * norm1), F(wpre1 * norm1), F(norm1)
    absval = F(0.5) * (np.abs(l0) + np.abs(l1))
    d = n20 / (n1 + absval) * (l0 - l1)
    s2 = n21 / (n1 + absval)
    small, big = F(0.55) * np.abs(l0), F(1.05) * np.abs(l0)
    neg = l0 < 0
    d = np.where(neg & (l1 > -small), d - s2 * (l1 + small), d)
    d = np.where(neg & ~(l1 > -small) & (l1 < -big), d + s2 * (-l1 - big), d)
    d = np.where(~neg & (l1 < small), d + s2 * (small - l1), d)
    d = np.where(~neg & ~(l1 < small) & (l1 > big), d - s2 * (l1 - big), d)
    d = d.astype(np.float32)
    h, w = d.shape
    q = np.pad(d.astype(np.float64), 4)
    acc = np.zeros((h, w))
    for pat in (_MALTA_LF if lf_patterns else _MALTA_HF):
        s = sum(q[4 + dy:4 + dy + h, 4 + dx:4 + dx + w] for dy, dx in pat)
        acc += s * s
    return acc.astype(np.float32)


def _l2_asym(a, b, w0gt1, w0lt1):
    vw0, vw1 = F(0.8 * w0gt1), F(0.8 * w0lt1)
    t = vw0 * (a - b) ** 2
    small, big = F(0.4) * np.abs(a), np.abs(a)
    neg = a < 0
    v = np.where(neg, np.where(b > -small, b + small, np.where(b < -big, -b - big, F(0))),
                 np.where(b < small, small - b, np.where(b > big, b - big, F(0))))
    return (t + vw1 * v * v).astype(np.float32)


def _fuzzy_erosion(bl):
    h, w = bl.shape
    inf = np.float32(np.inf)
    q = np.pad(bl, 3, constant_values=inf)
    cands = [bl, F(2.0) * bl, F(2.0) * bl]
    for dy in (-3, 0, 3):
        for dx in (-3, 0, 3):
            if dy or dx:
                cands.append(q[3 + dy:3 + dy + h, 3 + dx:3 + dx + w])
    st = np.sort(np.stack(cands), axis=0)
    return (F(0.45) * st[0] + F(0.3) * st[1] + F(0.25) * st[2]).astype(np.float32)


def _ba_diffmap(l1, l2, intensity, hf_asym=1.0, xmul=1.0):
    p0, p1 = _ba_psycho(l1, intensity), _ba_psycho(l2, intensity)
    ac = [np.zeros_like(l1[0]) for _ in range(3)]
    sq = math.sqrt(hf_asym)
    ac[1] += _malta(p0["uhf"][1], p1["uhf"][1], 1.10039032555 * hf_asym, 1.10039032555 / hf_asym, 71.7800275169, False)
    ac[0] += _malta(p0["uhf"][0], p1["uhf"][0], 173.5 * hf_asym, 173.5 / hf_asym, 5.0, False)
    ac[1] += _malta(p0["hf"][1], p1["hf"][1], 18.7237414387 * sq, 18.7237414387 / sq, 4498534.45232, True)
    ac[0] += _malta(p0["hf"][0], p1["hf"][0], 6923.99476109 * sq, 6923.99476109 / sq, 8051.15833247, True)
    ac[1] += _malta(p0["mf"][1], p1["mf"][1], 37.0819870399, 37.0819870399, 130262059.556, True)
    ac[0] += _malta(p0["mf"][0], p1["mf"][0], 8246.75321353, 8246.75321353, 1009002.70582, True)
    wmul = [400.0, 1.50815703118, 0, 2150.0, 10.6195433239, 16.2176043152, 29.2353797994, 0.844626970982, 0.703646627719]
    for c in range(2):
        ac[c] += _l2_asym(p0["hf"][c], p1["hf"][c], wmul[c] * hf_asym, wmul[c] / hf_asym)
    dc = []
    for c in range(3):
        ac[c] += F(wmul[3 + c]) * (p0["mf"][c] - p1["mf"][c]) ** 2
        dc.append(F(wmul[6 + c]) * (p0["lf"][c] - p1["lf"][c]) ** 2)
    mask = _fuzzy_erosion(p0["bl"])
    ac[1] += F(10.0) * (p0["bl"] - p1["bl"]) ** 2
    gs = F(1.0 / 17.83)
    mask_y = (gs * (F(1.0) + F(2.5485944793) / (F(0.451936922203) * mask + F(0.829591754942)))) ** 2
    mask_dc = (gs * (F(1.0) + F(0.505054525019) / (F(3.87449418804) * mask + F(0.20025578522)))) ** 2
    return np.sqrt(mask_dc * (F(xmul) * dc[0] + dc[1] + dc[2]) + mask_y * (F(xmul) * ac[0] + ac[1] + ac[2])).astype(np.float32)


def _ba_subsample(lin):
    _, h, w = lin.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = np.zeros((3, oh, ow), np.float32)
    for dy in range(2):
        for dx in range(2):
            part = lin[:, dy::2, dx::2]
            out[:, :part.shape[1], :part.shape[2]] += F(0.25) * part
    if w & 1:
        out[:, :, -1] *= F(2.0)
    if h & 1:
        out[:, -1, :] *= F(2.0)
    return out


def butteraugli(ref_u8: np.ndarray, dist_u8: np.ndarray, intensity: float = 80.0):
    """[h,w,3] uint8 pair -> (max of the diffmap, libjxl 3-norm) (A.5)."""
    l1, l2 = srgb8_to_linear(ref_u8), srgb8_to_linear(dist_u8)
    dm = _ba_diffmap(l1, l2, intensity)
    _, h, w = l1.shape
    if (w + 1) // 2 >= 8 and (h + 1) // 2 >= 8:
        ds = _ba_diffmap(_ba_subsample(l1), _ba_subsample(l2), intensity)
        up = ds[np.arange(h)[:, None] // 2, np.arange(w)[None, :] // 2]
        dm = dm * F(1.0 - 0.3 * 0.5) + F(0.5) * up
    d = dm.astype(np.float64)
    n = d.size
    pn = ((d ** 3).sum() / n) ** (1 / 3) + ((d ** 6).sum() / n) ** (1 / 6) + ((d ** 12).sum() / n) ** (1 / 12)
    return float(dm.max()), pn / 3.0
