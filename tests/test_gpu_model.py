"""GPU parity of the conv path through the C ABI vs the PyTorch fp32 oracle on identical seeded synthetic inputs and
weights.  Tolerances (north_star): fp32 parity mode max-abs <= 1e-3 on mask logits, Dice <= 1e-4; bf16 throughput
mode is checked against looser, explicitly stated bounds (1e-3 is not reachable with bf16 storage)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_LOGITS_FP32 = 1e-3
TOL_DICE = 1e-4
TOL_LOGITS_BF16 = 0.25          # absolute, logits calibrated to std 1.5
TOL_FLIP_BF16 = 0.02            # fraction of mask pixels allowed to differ in bf16 mode
TOL_DICE_BF16 = 5e-3


@pytest.fixture(scope="module")
def ysp():
    import yolo_u_b200
    assert torch.cuda.is_available()
    return yolo_u_b200


@pytest.fixture(scope="module")
def ref240(models):
    from oracle.model import synth_inputs, pipeline, mask_counts
    from oracle import nms as onms
    pred, seg = models
    x, lg, tg = synth_inputs(4, 240)
    with torch.no_grad():
        out = seg(x, lg)
        p_out, dets, keep, y, bott = pipeline(pred, seg, x, onms.non_max_suppression)
    return dict(x=x, lg=lg, tg=tg, seg_out=out, pipe_out=p_out, dets=dets, keep=keep, y=y, bott=bott,
                counts=mask_counts(p_out, tg))


def test_normalize_u8(ysp):
    import ctypes
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (3, 240, 240, 4), dtype=torch.uint8, generator=g)
    want = u8.permute(0, 3, 1, 2).float().div(255)               # ToTensor (dataset.py:68)
    d = u8.cuda()
    out = torch.empty(3, 4, 240, 240, device="cuda")
    rc = ysp.lib().ysp_normalize_u8(d.data_ptr(), out.data_ptr(), 3, 240, 240, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), want)


def test_mask_dice_counts(ysp):
    from oracle.model import mask_counts as omc
    g = torch.Generator().manual_seed(3)
    lg = torch.randn(5, 1, 240, 240, generator=g)
    lg[0, 0, 0, :4] = torch.tensor([0.0, 5e-8, -5e-8, 1e-6])     # sigmoid == 0.5 band
    tg = (torch.rand(5, 1, 240, 240, generator=g) > 0.5).float()
    tg[3] = 0
    lg[4] = -3.0
    counts, mask = ysp.mask_counts(lg.cuda(), tg.cuda(), want_mask=True)
    assert torch.equal(counts.cpu().long(), omc(lg, tg))
    assert torch.equal(mask.cpu().bool(), torch.sigmoid(lg) > 0.5)
    c2 = ysp.mask_counts(lg.cuda(), None)
    assert torch.equal(c2[:, 1].cpu(), counts[:, 1].cpu()) and int(c2[:, 0].sum()) == 0


PARITY_MODES = ["fp32", "tc32"]      # CUDA-core FFMA / tcgen05 with fp16 hi-lo operand splits: both must meet 1e-3


@pytest.mark.parametrize("mode", PARITY_MODES)
def test_segpp_fp32_parity(ysp, models, ref240, mode):
    pred, seg = models
    m = ysp.YOLOSegPlusPlus(pred, mode=mode)
    m.load_state_dict(seg.state_dict())
    out = m(ref240["x"].cuda(), ref240["lg"].cuda())
    assert out.shape == (4, 1, 240, 240) and out.dtype == torch.float32
    err = (out.cpu() - ref240["seg_out"]).abs().max().item()
    assert err <= TOL_LOGITS_FP32, f"max-abs logits error {err}"
    # reference-native resolution (160) and a non-square multiple of 8
    from oracle.model import synth_inputs
    for (h, w) in ((160, 160), (64, 96)):
        g = torch.Generator().manual_seed(h + w)
        x = torch.rand(2, 4, h, w, generator=g)
        lg = torch.sigmoid(torch.randn(2, 1, h // 8, w // 8, generator=g))
        with torch.no_grad():
            want = seg(x, lg)
        got = m(x.cuda(), lg.cuda())
        assert (got.cpu() - want).abs().max().item() <= TOL_LOGITS_FP32, (h, w)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 4, 100, 100).cuda(), torch.rand(1, 1, 12, 12).cuda())


@pytest.mark.parametrize("mode", PARITY_MODES)
def test_detector_fp32_parity(ysp, models, ref240, mode):
    pred, _ = models
    det = ysp.B200Detector.from_predictor(pred, mode=mode)
    from oracle.model import pad_to_multiple
    with torch.no_grad():
        y_ref, raws_ref = pred.model(pad_to_multiple(ref240["x"]))
    y, raws = det(ref240["x"].cuda())
    assert y.shape == (4, 5, 1344)
    # north_star bounds the MASK logits (1e-3) and Dice; the detector's raw head maps (logits of std 2, range +-15, computed
    # from features up to +-40) are bounded here at 1e-3 for the FFMA mode and 2.5e-3 for the fp16-split tensor-core mode
    tol_raw, tol_box = (1e-3, 2e-2) if mode == "fp32" else (2.5e-3, 6e-2)
    for r, rr in zip(raws, raws_ref):
        assert r.shape == rr.shape
        assert (r.cpu() - rr).abs().max().item() <= tol_raw
    assert (y[:, 4:].cpu() - y_ref[:, 4:]).abs().max().item() <= 2e-4         # sigmoid cls
    assert (y[:, :4].cpu() - y_ref[:, :4]).abs().max().item() <= tol_box      # boxes in pixels (x stride 8..32)
    # native %32 input, reference resolution
    x160 = torch.rand(2, 4, 160, 160, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        y_ref, raws_ref = pred.model(x160)
    y, raws = det(x160.cuda())
    assert y.shape == (2, 5, 525)
    assert (raws[0].cpu() - raws_ref[0]).abs().max().item() <= tol_raw


@pytest.mark.parametrize("mode", PARITY_MODES)
def test_pipeline_fp32_parity(ysp, models, ref240, mode):
    from oracle.model import dice_from_counts
    pred, seg = models
    P = ysp.Predictor.from_modules(pred, seg, mode=mode)
    ml, dets, keep, counts = P.predict(ref240["x"].cuda(), ref240["tg"].cuda())
    assert (ml.cpu() - ref240["pipe_out"]).abs().max().item() <= TOL_LOGITS_FP32
    d_ref = dice_from_counts(ref240["counts"])
    d = ysp.dice_from_counts(counts.cpu())
    assert (d - d_ref).abs().max().item() <= TOL_DICE
    # NMS inside the pipeline: bit-exactness is defined at the NMS boundary (same prediction tensor into both, SURVEY
    # App. C); the pipeline's own y differs from the oracle's by fp32 rounding, so compare against NMS of OUR y.
    o = P.predict_raw(ref240["x"].cuda(), ref240["tg"].cuda())
    from oracle import cnms
    want_d, want_k = cnms.nms_batched(o["y"].cpu(), 0.25, 0.45, 300)
    for b in range(4):
        assert torch.equal(keep[b].cpu(), want_k[b])
        assert torch.equal(dets[b].cpu(), want_d[b])
    # a3: sigmoid of the P3 class logit.  FFMA mode 1e-4; the fp16-split tensor-core mode carries the detector's raw-map
    # error (<= 2.5e-3, test_detector_fp32_parity) through sigmoid' <= 1/4 -- measured 1.2e-4; the bound north_star states
    # (mask logits 1e-3, Dice 1e-4) is asserted above and includes this input.
    assert (o["bottleneck"].cpu() - ref240["bott"]).abs().max().item() <= (1e-4 if mode == "fp32" else 3e-4)
    # u8 ingest path (a1) gives the same result as fp32 input of x/255
    u8 = (ref240["x"] * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    xf = u8.permute(0, 3, 1, 2).float() / 255
    a = P.predict_raw(u8.cuda())["mask_logits"].clone()
    b = P.predict_raw(xf.cuda())["mask_logits"].clone()
    assert torch.equal(a, b)


def test_u8_target_bitmask_and_large_max_det(ysp, models, ref240):
    """ysp_pipeline_io.d_target_u8 / d_mask_bits (the e2e path's formats) and max_det > 300 (workspace sized per call)."""
    pred, seg = models
    P = ysp.Predictor.from_modules(pred, seg, mode="tc32")
    x, tg = ref240["x"].cuda(), ref240["tg"].cuda()
    a = {k: v.clone() for k, v in P.predict_raw(x, tg).items()}
    tg_u8 = (ref240["tg"][:, 0] * 255).to(torch.uint8).cuda()                   # PNG bytes: 0 / 255
    b = P.predict_raw(x, tg_u8, want_bits=True)
    assert torch.equal(a["counts"], b["counts"]) and torch.equal(a["mask_logits"], b["mask_logits"])
    bits = b["mask_bits"].cpu().view(4, -1, 1)                                  # [B, HW/32, 1] int32
    unpacked = ((bits >> torch.arange(32, dtype=torch.int32)) & 1).view(4, 240, 240).bool()
    assert torch.equal(unpacked, torch.sigmoid(b["mask_logits"].cpu()[:, 0]) > 0.5)
    tg_u8[0, 0, :5] = torch.tensor([127, 128, 1, 254, 129], dtype=torch.uint8)  # v/255 > 0.5 <=> v >= 128
    c = P.predict_raw(x, tg_u8)["counts"].cpu()
    want_t = (tg_u8.cpu().float() / 255 > 0.5).view(4, -1).sum(1)
    assert torch.equal(c[:, 2].long(), want_t)
    ml, dets, keep, counts = P.predict(x, tg, conf_thres=0.001, iou_thres=0.7, max_det=1000)
    from oracle import cnms
    want_d, want_k = cnms.nms_batched(P._out["y"].cpu(), 0.001, 0.7, 1000)
    for i in range(4):
        assert torch.equal(keep[i].cpu(), want_k[i])


# bf16 (throughput mode, outside north_star's tolerance by construction: bf16 storage flips ~1 % of the mask pixels of these
# near-zero-mean logits): against a RANDOM target a 1 % flip rate moves a slice's Dice by up to ~1e-2 -- measured 7.9e-3 over
# the first 32 slices of the bench batch.  The parity mode (tc32) is held to north_star's 1e-3 / 1e-4.
@pytest.mark.parametrize("mode,tol_l,tol_d", [("tc32", TOL_LOGITS_FP32, TOL_DICE), ("bf16", 0.35, 1.5e-2)])
def test_parity_at_bench_batch(ysp, models, mode, tol_l, tol_d):
    """The configuration bench.py times: B = 256 slices per call (persistent-CTA tile scheduling differs from small
    batches).  The first 24 slices of the batch are checked against the CPU oracle."""
    from oracle.model import dice_from_counts, mask_counts, pipeline
    from oracle import nms as onms
    from oracle import cnms
    pred, seg = models
    g = torch.Generator().manual_seed(77)
    x = torch.rand(256, 4, 240, 240, generator=g)
    tg = (torch.rand(256, 1, 240, 240, generator=g) > 0.5).float()
    n = 24
    with torch.no_grad():
        want, _, _, _, _ = pipeline(pred, seg, x[:n], onms.non_max_suppression)
    P = ysp.Predictor.from_modules(pred, seg, mode=mode)
    o = P.predict_raw(x.cuda(), tg.cuda())
    torch.cuda.synchronize()
    ml = o["mask_logits"][:n].cpu()
    err = (ml - want).abs().max().item()
    d = ysp.dice_from_counts(o["counts"][:n].cpu())
    derr = (d - dice_from_counts(mask_counts(want, tg[:n]))).abs().max().item()
    print(f"{mode} @B=256: logits max-abs {err:.3e}, dice err {derr:.2e}")
    assert err <= tol_l and derr <= tol_d
    _, want_k = cnms.nms_batched(o["y"].cpu(), 0.25, 0.45, 300)
    cnt = o["det_count"].tolist()
    for b in range(256):
        assert torch.equal(o["det_idx"][b, :cnt[b]].cpu(), want_k[b])
    # the last slices of the batch agree with a small-batch run of the same slices (no tile-scheduling dependence)
    if mode == "tc32":
        big = o["mask_logits"][248:].clone()
        small = P.predict_raw(x[248:].cuda(), tg[248:].cuda())["mask_logits"]
        assert (small - big).abs().max().item() <= 2e-4


def test_bf16_mode_bounds(ysp, models, ref240):
    from oracle.model import dice_from_counts, mask_counts
    pred, seg = models
    P = ysp.Predictor.from_modules(pred, seg, mode="bf16")
    ml, dets, keep, counts = P.predict(ref240["x"].cuda(), ref240["tg"].cuda())
    ml = ml.cpu()
    err = (ml - ref240["pipe_out"]).abs().max().item()
    flips = ((ml > 0) != (ref240["pipe_out"] > 0)).float().mean().item()
    d = ysp.dice_from_counts(counts.cpu())
    d_ref = dice_from_counts(ref240["counts"])
    print(f"bf16: logits max-abs {err:.4f}, mask flips {flips:.5f}, dice err {(d - d_ref).abs().max().item():.2e}")
    assert err <= TOL_LOGITS_BF16
    assert flips <= TOL_FLIP_BF16
    assert (d - d_ref).abs().max().item() <= TOL_DICE_BF16
    assert torch.equal(counts.cpu().long(), mask_counts(ml, ref240["tg"]))   # counters exact on our own logits


def test_no_logit_ablation_variant():
    """`_YOLOSegPlusPlus.py` (reference ablation: decoder.0 = C3Ghost(128, 96), input = skip only): the engine detects
    it from the checkpoint's decoder.0.0.cv1 shape; parity vs the same modification of the oracle module."""
    import torch
    import yolo_u_b200 as ysp
    from oracle.model import build_models, synth_init_, synth_inputs
    from oracle.modules import C3Ghost
    pred, seg = build_models(0)
    seg.decoder[0][0] = C3Ghost(128, 96, n=1)
    synth_init_(seg.decoder[0][0], 11, lin_gain=2.0)
    seg.eval()

    def ref_forward(x):
        skips = []
        for idx, m in enumerate(seg.encoder):
            x = m(x)
            if idx in (2, 4):
                skips.append(x)
        for idx, m in enumerate(seg.decoder):
            if idx == 0:
                x = skips.pop()
            elif idx == 2:
                x = torch.cat([x, skips.pop()], 1)
            x = m(x)
        return seg.output(x)

    x, lg, _ = synth_inputs(2, 96)
    with torch.no_grad():
        want = ref_forward(x)
    m = ysp.YOLOSegPlusPlus(pred, mode="fp32", use_logits=False)
    m.load_state_dict(seg.state_dict(), strict=True)
    got = m(x.cuda())
    assert (got.cpu() - want).abs().max().item() <= 1e-3
    got2 = m(x.cuda(), lg.cuda())                       # a logits argument is accepted and ignored
    assert torch.equal(got, got2)
    mb = ysp.YOLOSegPlusPlus(pred, mode="bf16", use_logits=False)
    mb.load_state_dict(seg.state_dict(), strict=True)
    # bf16 storage: the synthetic head of this variant is not re-calibrated, so bound the error relative to the logit scale
    assert (mb(x.cuda()).cpu() - want).abs().max().item() <= 0.05 * want.abs().max().item() + 0.05
