"""Ingest oracle (oracle/ingest.py) pinned against the REAL cv2.resize outputs in tests/golden/ingest_golden.npz
(made by tests/golden/make_ingest_golden.py with cv2 4.13.0 in the build container)."""
import os

import numpy as np

from oracle import ingest as oi
from tests.golden.make_ingest_golden import inputs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest_golden.npz"))


def test_oracle_resize_is_bit_exact_with_cv2_golden():
    for i, (sh, sw, dh, dw) in enumerate(G["cases"].tolist()):
        img, mask = inputs(i, sh, sw)
        assert np.array_equal(oi.resize_linear_u8(img, dh, dw), G[f"lin{i}"]), (i, sh, sw, dh, dw)
        assert np.array_equal(oi.resize_nearest_u8(mask, dh, dw), G[f"nn{i}"]), (i, sh, sw, dh, dw)


def test_identity_and_area_special_case():
    img, _ = inputs(99, 64, 64)
    assert np.array_equal(oi.resize_linear_u8(img, 64, 64), img)
    down = oi.resize_linear_u8(img, 32, 32)
    s = img.astype(np.int64)
    assert np.array_equal(down, ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8))


def test_to_tensor_layout_and_scale():
    img, mask = inputs(7, 8, 6)
    t, m = oi.ingest(img, mask, 8)
    assert t.shape == (4, 8, 8) and m.shape == (1, 8, 8) and t.dtype == np.float32
    assert set(np.unique(m)).issubset({0.0, 1.0})
    assert float(t.max()) <= 1.0
