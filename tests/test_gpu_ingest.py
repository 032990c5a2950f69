"""Device ingest (ysp_resize_u8: cv2.resize + ToTensor) -- bit-exact against the cv2 golden vectors and the
numpy oracle, batched, plus the full-size case of the hot path (240 -> 240 identity, 155x240 -> 240, 480 -> 240)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest_golden.npz"))


def test_resize_matches_cv2_golden_bit_exact():
    import yolo_u_b200 as ysp
    from tests.golden.make_ingest_golden import inputs
    for i, (sh, sw, dh, dw) in enumerate(G["cases"].tolist()):
        img, mask = inputs(i, sh, sw)
        out = ysp.resize_u8(torch.from_numpy(img)[None].cuda(), (dh, dw), ysp.formats.INTER_LINEAR)
        assert np.array_equal(out[0].cpu().numpy(), G[f"lin{i}"]), (i, sh, sw, dh, dw)
        outm = ysp.resize_u8(torch.from_numpy(mask)[None].cuda(), (dh, dw), ysp.formats.INTER_NEAREST)
        assert np.array_equal(outm[0].cpu().numpy(), G[f"nn{i}"]), (i, sh, sw, dh, dw)


@pytest.mark.parametrize("sh,sw,S", [(240, 240, 240), (155, 240, 240), (480, 480, 240), (512, 500, 256), (200, 180, 160)])
def test_batched_ingest_matches_oracle(sh, sw, S):
    import yolo_u_b200 as ysp
    from oracle import ingest as oi
    rng = np.random.default_rng(sh * 7 + sw)
    B = 5
    img = rng.integers(0, 256, (B, sh, sw, 4), dtype=np.uint8)
    mask = (rng.integers(0, 2, (B, sh, sw), dtype=np.uint8) * 255).astype(np.uint8)
    t, m = ysp.ingest(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda(), S)
    u8 = ysp.resize_u8(torch.from_numpy(img).cuda(), S)
    for b in range(B):
        want_t, want_m = oi.ingest(img[b], mask[b], S)
        assert np.array_equal(t[b].cpu().numpy(), want_t)          # float32 value/255: same division, bit-exact
        assert np.array_equal(m[b].cpu().numpy(), want_m)
        assert np.array_equal(u8[b].cpu().numpy(), oi.resize_linear_u8(img[b], S, S))


def test_ingest_feeds_the_pipeline_u8_input():
    """resized uint8 [B,S,S,4] goes straight into the pipeline's uint8 path (the stem normalises on load)."""
    import yolo_u_b200 as ysp
    from oracle.model import build_models
    pred, seg = build_models(0)
    P = ysp.Predictor.from_modules(pred, seg, device="cuda:0", mode="fp32")
    rng = np.random.default_rng(3)
    raw = torch.from_numpy(rng.integers(0, 256, (2, 300, 280, 4), dtype=np.uint8)).cuda()
    u8 = ysp.resize_u8(raw, 240)
    f32 = ysp.resize_u8(raw, 240, to_tensor=True)
    a = P.predict_raw(u8)["mask_logits"].clone()
    b = P.predict_raw(f32)["mask_logits"].clone()
    assert torch.equal(a, b)


def test_errors():
    import yolo_u_b200 as ysp
    with pytest.raises(TypeError):
        ysp.resize_u8(torch.zeros(1, 8, 8, 4, device="cuda"), 8)
    with pytest.raises(ValueError):
        ysp.resize_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device="cuda"), 8)
    with pytest.raises(Exception):
        ysp.resize_u8(torch.zeros(1, 8, 8, 4, dtype=torch.uint8), 8)
