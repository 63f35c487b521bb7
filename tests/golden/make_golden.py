"""Generates tests/golden/pairs.npz: small seeded reference/distorted pairs and the ORACLE's outputs for them.

The reference implementation itself cannot run in this container (no Rust toolchain, un-vendored crates:
SURVEY.md 8c), so these vectors pin the oracle (oracle/ce_oracle.c) against regressions and give the GPU
tests committed inputs; they are not outputs of the reference.  Re-run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from codec_eval_b200.synth import G, J  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [  # (seed, w, h, quality, subsampling)
    (0, 64, 64, 80, 2),
    (1, 96, 80, 50, 2),
    (2, 127, 61, 90, 0),   # odd sizes: exercises every downsample edge rule
    (3, 256, 256, 75, 2),
    (4, 40, 24, 30, 2),
    (5, 512, 512, 80, 2),  # BASELINE.json configs[0]
]


def main():
    out = {}
    rows = []
    for k, (seed, w, h, q, ss) in enumerate(CASES):
        ref = G(seed, w, h)
        dist = J(ref, q, ss)
        out[f"ref{k}"] = ref
        out[f"dist{k}"] = dist
        s2 = O.ssimulacra2(ref, dist, w, h)
        ds = O.dssim(ref, dist, w, h)
        ba, pn = O.butteraugli(ref, dist, w, h)
        sse = O.sse(ref, dist)
        ps = O.psnr(ref, dist, w, h)
        rows.append([seed, w, h, q, ss, sse, ps, s2, ds, ba, pn])
        print(rows[-1])
    out["table"] = np.array(rows, dtype=np.float64)
    out["xyb_rt0"] = O.xyb_roundtrip(out["ref0"], 64, 64).reshape(64, 64, 3)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "pairs.npz"), **out)


if __name__ == "__main__":
    main()
