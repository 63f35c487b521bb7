"""Minimal ICC v2 / v4 matrix/TRC RGB profiles built byte by byte for the ICC tests (no profile files are shipped)."""
import struct

import numpy as np

D50 = (0.9642, 1.0, 0.8249)


def _s15(v):
    return struct.pack(">i", int(round(v * 65536.0)))


def _xyz_tag(xyz):
    return b"XYZ " + b"\0" * 4 + b"".join(_s15(v) for v in xyz)


def curv_gamma(g):
    return b"curv" + b"\0" * 4 + struct.pack(">IH", 1, int(round(g * 256))) + b"\0\0"


def curv_table(values):
    v = np.clip(np.round(np.asarray(values) * 65535.0), 0, 65535).astype(">u2")
    return b"curv" + b"\0" * 4 + struct.pack(">I", len(v)) + v.tobytes() + (b"\0\0" if len(v) % 2 else b"")


def para(ftype, params):
    return b"para" + b"\0" * 4 + struct.pack(">HH", ftype, 0) + b"".join(_s15(p) for p in params)


SRGB_PARA = para(3, [2.4, 1 / 1.055, 0.055 / 1.055, 1 / 12.92, 0.04045])


def _bradford_to_d50(white_xy):
    """Bradford adaptation matrix from the white point (x, y) to D50."""
    wx, wy = white_xy
    W = np.array([wx / wy, 1.0, (1 - wx - wy) / wy])
    B = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]])
    s, d = B @ W, B @ np.array(D50)
    return np.linalg.inv(B) @ np.diag(d / s) @ B


def colorants(primaries_xy, white_xy):
    """3x3 matrix whose columns are the D50-adapted XYZ of the R, G, B primaries."""
    P = np.array([[x / y, 1.0, (1 - x - y) / y] for x, y in primaries_xy]).T
    wx, wy = white_xy
    W = np.array([wx / wy, 1.0, (1 - wx - wy) / wy])
    S = np.linalg.solve(P, W)
    return _bradford_to_d50(white_xy) @ (P * S)


def make_profile(primaries_xy, white_xy, trc, version=0x02400000, space=b"RGB ", pcs=b"XYZ ", drop=()):
    """trc: one tag body for all channels, or a list of three."""
    M = colorants(primaries_xy, white_xy)
    trcs = trc if isinstance(trc, (list, tuple)) else [trc] * 3
    desc = b"desc" + b"\0" * 4 + struct.pack(">I", 5) + b"test\0" + b"\0" * (4 + 4 + 2 + 1 + 67)
    tags = [(b"desc", desc), (b"wtpt", _xyz_tag(D50)), (b"cprt", b"text" + b"\0" * 4 + b"none\0")]
    for i, sig in enumerate((b"rXYZ", b"gXYZ", b"bXYZ")):
        tags.append((sig, _xyz_tag(M[:, i])))
    for i, sig in enumerate((b"rTRC", b"gTRC", b"bTRC")):
        tags.append((sig, trcs[i]))
    tags = [t for t in tags if t[0] not in drop]
    body, table, off = b"", b"", 128 + 4 + 12 * len(tags)
    for sig, data in tags:
        pad = (-len(data)) % 4
        table += sig + struct.pack(">II", off + len(body), len(data))
        body += data + b"\0" * pad
    size = 128 + 4 + len(table) + len(body)
    hdr = struct.pack(">I4sI4s4s4s", size, b"lcms", version, b"mntr", space, pcs) + b"\0" * 12 + b"acsp" + b"APPL" + b"\0" * 4 + \
        b"\0" * 4 + b"\0" * 4 + b"\0" * 8 + b"\0" * 4 + b"".join(_s15(v) for v in D50) + b"lcms" + b"\0" * 16 + b"\0" * 28
    assert len(hdr) == 128
    return hdr + struct.pack(">I", len(tags)) + table + body


P3 = [(0.680, 0.320), (0.265, 0.690), (0.150, 0.060)]
ADOBE = [(0.64, 0.33), (0.21, 0.71), (0.15, 0.06)]
REC2020 = [(0.708, 0.292), (0.170, 0.797), (0.131, 0.046)]
SRGB = [(0.64, 0.33), (0.30, 0.60), (0.15, 0.06)]
D65 = (0.3127, 0.3290)


def display_p3():
    return make_profile(P3, D65, SRGB_PARA, version=0x04000000)


def adobe_rgb():
    return make_profile(ADOBE, D65, curv_gamma(563 / 256.0))


def rec2020_table():
    x = np.linspace(0, 1, 1024)
    return make_profile(REC2020, D65, curv_table(x ** 2.4))


def srgb_like():
    return make_profile(SRGB, D65, SRGB_PARA, version=0x04000000)
