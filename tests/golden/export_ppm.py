"""Writes the golden pairs of tests/golden/pairs.npz as binary PPM files (case<k>_ref.ppm / case<k>_dist.ppm) for
rust/pin-parity, which runs the REAL codec-eval metric functions on them (see that crate's Cargo.toml).

  python tests/golden/export_ppm.py /tmp/pin
"""
import os
import sys

import numpy as np


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pairs.npz"))
    k = 0
    while f"ref{k}" in z.files:
        for side in ("ref", "dist"):
            img = z[f"{side}{k}"]
            h, w, _ = img.shape
            with open(os.path.join(out_dir, f"case{k}_{side}.ppm"), "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (w, h))
                f.write(np.ascontiguousarray(img, np.uint8).tobytes())
        k += 1
    print(f"wrote {k} pairs to {out_dir}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pin_ppm")
