"""Writes tests/golden/jpeg_golden.npz: inputs and the decoded output of a REAL baseline JPEG codec (Pillow's
libjpeg-turbo: save -> load) for a few small images.  Run in the build container:  python tests/golden/make_jpeg_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from codec_eval_b200.synth import G, J  # noqa: E402

cases = [(64, 48, 75, 2), (64, 48, 90, 0), (100, 60, 50, 2), (77, 35, 85, 2), (9, 33, 30, 2), (40, 24, 95, 0), (1, 1, 75, 2),
         (33, 50, 10, 2)]
out = {"table": np.asarray(cases, np.int32)}
for k, (w, h, q, ss) in enumerate(cases):
    src = G(100 + k, w, h)
    out[f"src{k}"] = src
    out[f"out{k}"] = J(src, q, ss)
import PIL  # noqa: E402

out["pillow_version"] = np.asarray(PIL.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "jpeg_golden.npz"), **out)
print("wrote", len(cases), "cases; Pillow", PIL.__version__)
