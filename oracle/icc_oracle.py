"""CPU oracle of the ICC matrix/TRC -> sRGB transform.  TEST INFRASTRUCTURE ONLY.

Restates, independently of codec_eval_b200/csrc/k_icc.cu, what src/metrics/icc.rs:69-103 asks of its CMS for an RGB
display profile: per-channel tone curve -> colorant matrix into the D50 PCS -> inverse sRGB colorants -> sRGB OETF ->
8 bits (ICC.1:2010 10.5 curveType, 10.15 parametricCurveType, Annex F.3 three-component matrix model).

PARITY UNPINNED against the reference's CMS (moxcms, crates.io, not vendored).  tests/test_icc.py pins this file against
another real CMS instead -- Little CMS 2 through Pillow's ImageCms -- to within 2 code values (CMS engines differ in
table sizes and fixed-point paths; the reference's own notes put moxcms and lcms2 1-2.5 SSIMULACRA2 points apart,
icc.rs:18-26).

Host tables are built with math.pow (the C library's pow, as the product's host code uses); the per-pixel arithmetic is
numpy float32, one IEEE operation at a time, in the kernel's order, so the CUDA path can match bit for bit.
"""
from __future__ import annotations

import math
import struct

import numpy as np

SRGB_D50 = ((0.4360747, 0.3850649, 0.1430804), (0.2225045, 0.7168786, 0.0606169), (0.0139322, 0.0971045, 0.7141733))


class IccError(Exception):
    pass


def _tags(icc: bytes):
    if len(icc) < 132:
        raise IccError("shorter than its header")
    if icc[36:40] != b"acsp":
        raise IccError("missing 'acsp' signature")
    if icc[16:20] != b"RGB ":
        raise IccError("not an RGB profile")
    if icc[20:24] != b"XYZ ":
        raise IccError("profile connection space is not XYZ")
    (count,) = struct.unpack(">I", icc[128:132])
    if 132 + count * 12 > len(icc):
        raise IccError("tag table exceeds the profile")
    tags = {}
    for i in range(count):
        sig, off, size = struct.unpack(">4sII", icc[132 + 12 * i:144 + 12 * i])
        if off + size > len(icc) or size < 8:
            raise IccError(f"tag {sig!r} out of bounds")
        tags.setdefault(sig, icc[off:off + size])
    return tags


def _s15f16(b: bytes) -> float:
    return struct.unpack(">i", b)[0] / 65536.0


def _trc(tag: bytes, x: float) -> float:
    kind = tag[:4]
    if kind == b"curv":
        (cnt,) = struct.unpack(">I", tag[8:12])
        if 12 + 2 * cnt > len(tag):
            raise IccError("truncated curve")
        if cnt == 0:
            return x
        if cnt == 1:
            return math.pow(x, struct.unpack(">H", tag[12:14])[0] / 256.0)
        pos = x * (cnt - 1)
        i0 = min(int(pos), cnt - 2)
        f = pos - i0
        a = struct.unpack(">H", tag[12 + 2 * i0:14 + 2 * i0])[0] / 65535.0
        b = struct.unpack(">H", tag[14 + 2 * i0:16 + 2 * i0])[0] / 65535.0
        return a + (b - a) * f
    if kind == b"para":
        (ft,) = struct.unpack(">H", tag[8:10])
        n = (1, 3, 4, 5, 7)[ft] if ft <= 4 else None
        if n is None or 12 + 4 * n > len(tag):
            raise IccError("bad parametric curve")
        q = [_s15f16(tag[12 + 4 * i:16 + 4 * i]) for i in range(n)] + [0.0] * (7 - n)
        g, a, b, c, d, e, f = q
        if ft == 0:
            return math.pow(x, g)
        if ft == 1:
            return math.pow(a * x + b, g) if x >= -b / a else 0.0
        if ft == 2:
            return math.pow(a * x + b, g) + c if x >= -b / a else c
        if ft == 3:
            return math.pow(a * x + b, g) if x >= d else c * x
        return math.pow(a * x + b, g) + e if x >= d else c * x + f
    raise IccError("unsupported curve type")


def _srgb_oetf(l: float) -> float:
    return 12.92 * l if l <= 0.0031308 else 1.055 * math.pow(l, 1.0 / 2.4) - 0.055


_OUT_LUT = None


def _out_lut() -> np.ndarray:
    global _OUT_LUT
    if _OUT_LUT is None:
        t = np.empty(65536, np.uint8)
        for i in range(65536):
            v = _srgb_oetf(i / 65535.0) * 255.0 + 0.5
            t[i] = 0 if v < 0.0 else (255 if v > 255.0 else int(math.floor(v)))
        _OUT_LUT = t
    return _OUT_LUT


def tables(icc: bytes):
    """-> (in_lut float32 [3,256], matrix float32 [3,3]) or raises IccError."""
    tags = _tags(icc)
    P = [[0.0] * 3 for _ in range(3)]
    for c, name in enumerate((b"rXYZ", b"gXYZ", b"bXYZ")):
        t = tags.get(name)
        if t is None or t[:4] != b"XYZ " or len(t) < 20:
            raise IccError(f"no usable {name.decode()} tag (not a matrix/TRC profile)")
        for r in range(3):
            P[r][c] = _s15f16(t[8 + 4 * r:12 + 4 * r])
    lut = np.empty((3, 256), np.float32)
    for c, name in enumerate((b"rTRC", b"gTRC", b"bTRC")):
        t = tags.get(name)
        if t is None:
            raise IccError(f"no {name.decode()} tag (not a matrix/TRC profile)")
        for i in range(256):
            lut[c, i] = np.float32(_trc(t, i / 255.0))
    S = SRGB_D50
    det = (S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0])
           + S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]))
    Si = [[(S[1][1] * S[2][2] - S[1][2] * S[2][1]) / det, (S[0][2] * S[2][1] - S[0][1] * S[2][2]) / det,
           (S[0][1] * S[1][2] - S[0][2] * S[1][1]) / det],
          [(S[1][2] * S[2][0] - S[1][0] * S[2][2]) / det, (S[0][0] * S[2][2] - S[0][2] * S[2][0]) / det,
           (S[0][2] * S[1][0] - S[0][0] * S[1][2]) / det],
          [(S[1][0] * S[2][1] - S[1][1] * S[2][0]) / det, (S[0][1] * S[2][0] - S[0][0] * S[2][1]) / det,
           (S[0][0] * S[1][1] - S[0][1] * S[1][0]) / det]]
    M = np.empty((3, 3), np.float32)
    for r in range(3):
        for c in range(3):
            M[r, c] = np.float32((Si[r][0] * P[0][c] + Si[r][1] * P[1][c]) + Si[r][2] * P[2][c])
    return lut, M


def transform_to_srgb(rgb: np.ndarray, icc: bytes | None) -> np.ndarray:
    """rgb uint8 [..., 3] -> uint8 of the same shape (src/metrics/icc.rs:69-103)."""
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    if not icc:
        return a.copy()
    lut, M = tables(icc)
    px = a.reshape(-1, 3)
    r, g, b = lut[0][px[:, 0]], lut[1][px[:, 1]], lut[2][px[:, 2]]
    out = np.empty_like(px)
    olut = _out_lut()
    for c in range(3):
        v = (M[c, 0] * r + M[c, 1] * g) + M[c, 2] * b          # float32, one rounding per operation
        v = np.minimum(np.maximum(v, np.float32(0.0)), np.float32(1.0))
        out[:, c] = olut[(v * np.float32(65535.0) + np.float32(0.5)).astype(np.int32)]
    return out.reshape(a.shape)
