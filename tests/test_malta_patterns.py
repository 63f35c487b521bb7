"""The CUDA Malta line sums share sub-sums between the 16 orientations (k_butteraugli.cu malta_hf / malta_lf).
This expands the C++ expressions symbolically and checks that every pattern is exactly the oracle's tap set
(oracle/ce_oracle.c MALTA_HF / MALTA_LF, i.e. libjxl MaltaUnit / MaltaUnitLF).  CPU only."""
import ctypes as C
import os
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Taps(Counter):
    def __add__(self, other):
        r = Taps(self)
        r.update(other)
        return r


def _expand(src, name):
    a = src.index(f"CE_DEVINL float {name}(")
    body = src[a:src.index("return acc;", a)]
    env = {"D": lambda dy, dx: Taps({(dy, dx): 1})}
    patterns = []
    for line in body.split("\n")[1:]:
        line = line.strip()
        if line.startswith("const float"):
            depth, cur, parts = 0, "", []
            for ch in line[len("const float"):].rstrip(";"):
                depth += ch == "("
                depth -= ch == ")"
                if ch == "," and depth == 0:
                    parts.append(cur)
                    cur = ""
                else:
                    cur += ch
            parts.append(cur)
            for p in parts:
                k, v = p.split("=", 1)
                env[k.strip()] = eval(v, env)
        elif line.startswith("t ="):
            patterns.append(eval(line[3:].rstrip(";"), env))
    return patterns


def test_cuda_malta_sums_are_the_oracle_patterns(O):
    src = open(os.path.join(ROOT, "codec_eval_b200", "csrc", "k_butteraugli.cu")).read()
    L = O.lib()
    hf = np.zeros((16, 9, 2), np.int8)
    hn = np.zeros(16, np.int8)
    lf = np.zeros((16, 5, 2), np.int8)
    L.ceo_malta_patterns(hf.ctypes.data_as(C.c_void_p), hn.ctypes.data_as(C.c_void_p), lf.ctypes.data_as(C.c_void_p))
    got = _expand(src, "malta_hf")
    assert len(got) == 16
    for p in range(16):
        exp = Counter((int(hf[p, t, 0]), int(hf[p, t, 1])) for t in range(int(hn[p])))
        assert got[p] == exp, (p, got[p], exp)
    got = _expand(src, "malta_lf")
    assert len(got) == 16
    for p in range(16):
        exp = Counter((int(lf[p, t, 0]), int(lf[p, t, 1])) for t in range(5))
        assert got[p] == exp, (p, got[p], exp)


def test_generated_shared_sums_are_the_oracle_patterns(O):
    """csrc/malta_sums.inc (tools/gen_malta.py): the sums of a thread's 4 x 2 pixels with sub-sums shared between lines
    and pixels.  Every `s` of pixel (J, K), line p must expand to exactly the oracle's taps of line p, shifted to the
    pixel's place in the 10 x 12 window (row = dy + 4 + J, column = dx + 4 + K), each tap once."""
    import re

    src = open(os.path.join(ROOT, "codec_eval_b200", "csrc", "malta_sums.inc")).read()
    L = O.lib()
    hf = np.zeros((16, 9, 2), np.int8)
    hn = np.zeros(16, np.int8)
    lf = np.zeros((16, 5, 2), np.int8)
    L.ceo_malta_patterns(hf.ctypes.data_as(C.c_void_p), hn.ctypes.data_as(C.c_void_p), lf.ctypes.data_as(C.c_void_p))
    tables = {"malta_hf8": [[(int(hf[p, t, 0]), int(hf[p, t, 1])) for t in range(int(hn[p]))] for p in range(16)],
              "malta_lf8": [[(int(lf[p, t, 0]), int(lf[p, t, 1])) for t in range(5)] for p in range(16)]}

    class Win:
        def __getitem__(self, r):
            return {c: Taps({(r, c): 1}) for c in range(12)}

    for fn, table in tables.items():
        a = src.index(f"CE_DEVINL void {fn}(")
        body = src[a:src.index("\n}", a)]
        env = {"win": Win()}
        seen = set()
        adds = 0
        for line in body.split("\n")[1:]:
            line = line.strip()
            m = re.match(r"const float (t\d+) = (.*);$", line)
            if m:
                env[m.group(1)] = eval(m.group(2), env)
                adds += m.group(2).count("+")
                continue
            m = re.match(r"\{ const float s = (.*); acc\[(\d)\]\[(\d)\] = __fmaf_rn\(s, s, acc\[\2\]\[\3\]\); \}\s+// pixel \((\d),(\d)\) line (\d+)$", line)
            assert m, line
            J, K, p = int(m.group(2)), int(m.group(3)), int(m.group(6))
            assert (J, K) == (int(m.group(4)), int(m.group(5)))
            got = eval(m.group(1), env)
            adds += m.group(1).count("+")
            exp = Counter((dy + 4 + J, dx + 4 + K) for dy, dx in table[p])
            assert got == exp, (fn, J, K, p)
            seen.add((J, K, p))
        assert seen == {(J, K, p) for J in range(2) for K in range(4) for p in range(16)}
        # the point of the sharing: fewer adds than one pixel at a time (90 / 48 per pixel in malta_hf / malta_lf)
        assert adds / 8 < (60 if fn == "malta_hf8" else 46), (fn, adds / 8)
