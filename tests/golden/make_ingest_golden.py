"""Golden vectors for the slice ingest: outputs of the REAL `cv2.resize` (the reference's own call,
/root/reference/dataset.py:59-65) on seeded uint8 inputs.  Run in the build container (cv2 4.13.0):
    python tests/golden/make_ingest_golden.py
Inputs are regenerated from the seed by the tests; only sizes + cv2 outputs are stored."""
import os

import numpy as np      # cv2 is imported in main() only: the tests import `inputs` from here and must not need OpenCV

CASES = [(155, 200, 96, 96), (96, 96, 96, 96), (75, 65, 60, 60), (60, 60, 120, 120),
         (128, 128, 64, 64), (128, 80, 64, 64), (97, 131, 64, 48), (64, 48, 97, 131), (33, 17, 128, 96), (1, 1, 8, 8),
         (119, 121, 120, 120), (240, 240, 160, 160)]


def inputs(i, sh, sw):
    rng = np.random.default_rng(1000 + i)
    img = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8)
    mask = (rng.integers(0, 2, (sh, sw), dtype=np.uint8) * 255).astype(np.uint8)
    return img, mask


def main():
    import cv2
    out = {"cases": np.array(CASES, dtype=np.int32), "cv2_version": np.array(cv2.__version__)}
    for i, (sh, sw, dh, dw) in enumerate(CASES):
        img, mask = inputs(i, sh, sw)
        out[f"lin{i}"] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out[f"nn{i}"] = cv2.resize(mask, (dw, dh), interpolation=cv2.INTER_NEAREST)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_golden.npz"), **out)


if __name__ == "__main__":
    main()
