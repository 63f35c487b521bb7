#!/usr/bin/env python
"""Generates codec_eval_b200/csrc/malta_sums.inc: the 16 oriented Malta line sums (libjxl MaltaUnit / MaltaUnitLF, the
tap tables of SURVEY.md A.5) of the 4 x 2 pixels a thread of k_ba_malta owns, as straight-line adds over its 10 x 12
register window, with the sub-sums that several lines / several of the 8 pixels have in common formed once.

The sharing is found by greedy common-pair elimination over the 128 sums (8 pixels x 16 lines): the pair of terms that
occurs together in the most sums becomes a temporary, until no pair occurs twice; ties are broken at random and the
best of `--tries` runs is kept.  Mathematically every line is the upstream sum; only the association differs.

  python tools/gen_malta.py [--tries 200] [--seed 1] [--check]      (--check: regenerate and compare with the file)
"""
import argparse
import os
import random
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "codec_eval_b200", "csrc", "malta_sums.inc")

HF = [
    [(0, -4), (0, -3), (0, -2), (0, -1), (0, 0), (0, 1), (0, 2), (0, 3), (0, 4)],
    [(-4, 0), (-3, 0), (-2, 0), (-1, 0), (0, 0), (1, 0), (2, 0), (3, 0), (4, 0)],
    [(-3, -3), (-2, -2), (-1, -1), (0, 0), (1, 1), (2, 2), (3, 3)],
    [(-3, 3), (-2, 2), (-1, 1), (0, 0), (1, -1), (2, -2), (3, -3)],
    [(-4, 1), (-3, 1), (-2, 1), (-1, 0), (0, 0), (1, 0), (2, -1), (3, -1), (4, -1)],
    [(-4, -1), (-3, -1), (-2, -1), (-1, 0), (0, 0), (1, 0), (2, 1), (3, 1), (4, 1)],
    [(-1, -4), (-1, -3), (-1, -2), (0, -1), (0, 0), (0, 1), (1, 2), (1, 3), (1, 4)],
    [(1, -4), (1, -3), (1, -2), (0, -1), (0, 0), (0, 1), (-1, 2), (-1, 3), (-1, 4)],
    [(-3, -2), (-2, -1), (-1, -1), (0, 0), (1, 1), (2, 1), (3, 2)],
    [(-3, 2), (-2, 1), (-1, 1), (0, 0), (1, -1), (2, -1), (3, -2)],
    [(-2, -3), (-1, -2), (-1, -1), (0, 0), (1, 1), (1, 2), (2, 3)],
    [(-2, 3), (-1, 2), (-1, 1), (0, 0), (1, -1), (1, -2), (2, -3)],
    [(2, -4), (2, -3), (1, -2), (1, -1), (0, 0), (0, 1), (-1, 2), (-1, 3)],
    [(-2, -4), (-2, -3), (-1, -2), (-1, -1), (0, 0), (0, 1), (1, 2), (1, 3)],
    [(-4, -2), (-3, -2), (-2, -1), (-1, -1), (0, 0), (1, 0), (2, 1), (3, 1)],
    [(-4, 2), (-3, 2), (-2, 1), (-1, 1), (0, 0), (1, 0), (2, -1), (3, -1)],
]
LF = [
    [(0, -4), (0, -2), (0, 0), (0, 2), (0, 4)],
    [(-4, 0), (-2, 0), (0, 0), (2, 0), (4, 0)],
    [(-3, -3), (-2, -2), (0, 0), (2, 2), (3, 3)],
    [(-3, 3), (-2, 2), (0, 0), (2, -2), (3, -3)],
    [(-4, 1), (-2, 1), (0, 0), (2, -1), (4, -1)],
    [(-4, -1), (-2, -1), (0, 0), (2, 1), (4, 1)],
    [(-1, -4), (-1, -2), (0, 0), (1, 2), (1, 4)],
    [(1, -4), (1, -2), (0, 0), (-1, 2), (-1, 4)],
    [(-3, -2), (-2, -1), (0, 0), (2, 1), (3, 2)],
    [(-3, 2), (-2, 1), (0, 0), (2, -1), (3, -2)],
    [(-2, -3), (-1, -2), (0, 0), (1, 2), (2, 3)],
    [(-2, 3), (-1, 2), (0, 0), (1, -2), (2, -3)],
    [(2, -4), (1, -2), (0, 0), (-1, 2), (-2, 4)],
    [(-2, -4), (-1, -2), (0, 0), (1, 2), (2, 4)],
    [(-4, -2), (-2, -1), (0, 0), (2, 1), (4, 2)],
    [(-4, 2), (-2, 1), (0, 0), (2, -1), (4, -2)],
]
ROWS, COLS = 2, 4   # pixels per thread: rows J, columns K


def sums_of(patterns):
    """[(J, K, p, frozenset of window cells)]; cell = ('w', row, col) with row = dy + 4 + J, col = dx + 4 + K."""
    out = []
    for J in range(ROWS):
        for K in range(COLS):
            for p, taps in enumerate(patterns):
                out.append((J, K, p, [("w", dy + 4 + J, dx + 4 + K) for dy, dx in taps]))
    return out


def eliminate(sums, rng):
    """Greedy common-pair elimination.  Returns (temps [(name, a, b)], reduced sums)."""
    sums = [list(s) for s in sums]
    temps = []
    while True:
        cnt = Counter()
        for s in sums:
            ss = sorted(s)
            for i in range(len(ss)):
                for j in range(i + 1, len(ss)):
                    cnt[(ss[i], ss[j])] += 1
        if not cnt:
            break
        best = max(cnt.values())
        if best < 2:
            break
        cands = [k for k, v in cnt.items() if v == best]
        a, b = rng.choice(cands)
        t = ("t", len(temps), 0)
        temps.append((t, a, b))
        for s in sums:
            if a in s and b in s:
                s.remove(a)
                s.remove(b)
                s.append(t)
    return temps, sums


def cost(temps, sums):
    return len(temps) + sum(len(s) - 1 for s in sums)


def name(sym):
    return f"win[{sym[1]}][{sym[2]}]" if sym[0] == "w" else f"t{sym[1]}"


def depth_of(sym, depth):
    return 0 if sym[0] == "w" else depth[sym]


def emit(fn, patterns, tries, seed):
    base = sums_of(patterns)
    best = None
    for i in range(tries):
        rng = random.Random(seed * 100003 + i)
        temps, red = eliminate([s[3] for s in base], rng)
        c = cost(temps, red)
        if best is None or c < best[0]:
            best = (c, temps, red)
    c, temps, red = best
    naive = sum(len(s[3]) - 1 for s in base)
    lines = [f"// {fn}: {c} adds for {ROWS * COLS} pixels x 16 lines ({c / (ROWS * COLS):.1f} per pixel; {naive // (ROWS * COLS)} without sharing)",
             f"CE_DEVINL void {fn}(const float (&win)[MT_WIN_ROWS][12], float (&acc)[{ROWS}][{COLS}]) {{"]
    depth = {}
    for t, a, b in temps:
        depth[t] = 1 + max(depth_of(a, depth), depth_of(b, depth))
        lines.append(f"    const float {name(t)} = {name(a)} + {name(b)};")
    # line sums: shallow terms first so the add tree stays short; squares accumulate per pixel in pattern order
    for (J, K, p, _), s in zip(base, red):
        terms = sorted(s, key=lambda x: (depth_of(x, depth), x))
        # balanced pairing
        exprs = [name(x) for x in terms]
        while len(exprs) > 1:
            nxt = [f"({exprs[i]} + {exprs[i + 1]})" for i in range(0, len(exprs) - 1, 2)]
            if len(exprs) & 1:
                nxt.append(exprs[-1])
            exprs = nxt
        e = exprs[0]
        if e.startswith("("):
            e = e[1:-1]
        lines.append(f"    {{ const float s = {e}; acc[{J}][{K}] = __fmaf_rn(s, s, acc[{J}][{K}]); }}   // pixel ({J},{K}) line {p}")
    lines.append("}")
    return "\n".join(lines), c


def generate(tries, seed):
    hf, chf = emit("malta_hf8", HF, tries, seed)
    lf, clf = emit("malta_lf8", LF, tries, seed)
    head = ("// GENERATED by tools/gen_malta.py -- do not edit.  The 16 Malta line sums of a thread's 4 x 2 pixels over its\n"
            "// 10 x 12 register window (row = dy + 4 + J, column = dx + 4 + K), sub-sums shared between lines and pixels.\n"
            f"// tries={tries} seed={seed}\n")
    return head + hf + "\n\n" + lf + "\n", chf, clf


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tries", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    text, chf, clf = generate(a.tries, a.seed)
    print(f"hf: {chf} adds / 8 px = {chf / 8:.2f} per pixel;  lf: {clf} adds / 8 px = {clf / 8:.2f} per pixel")
    if a.check:
        assert open(OUT).read() == text, "malta_sums.inc is stale"
    else:
        open(OUT, "w").write(text)
