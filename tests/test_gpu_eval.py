"""GPU: evaluation-loop level behaviour -- sharded volumes (BASELINE cfg 3 in miniature), the double-buffered host
pipeline, odd batch sizes, reference-native 160x160 resolution, error behaviour."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ysp():
    import yolo_u_b200
    assert torch.cuda.is_available()
    return yolo_u_b200


@pytest.fixture(scope="module")
def predictor(ysp, models):
    pred, seg = models
    return ysp.Predictor.from_modules(pred, seg, mode="fp32")


def _oracle_counts(models, x, tg):
    from oracle.model import pipeline, mask_counts
    from oracle import nms as onms
    pred, seg = models
    with torch.no_grad():
        out, dets, keep, y, bott = pipeline(pred, seg, x, onms.non_max_suppression)
    return out, mask_counts(out, tg)


def test_sharded_volumes_match_single_pass(ysp, models, predictor):
    """2 'volumes' x 5 slices, sharded over 2 ranks and batched by 3 (ragged last batch): the merged counters equal the
    oracle's over all 10 slices (Dice <= 1e-4, counters equal up to the fp32 1e-3 logit tolerance => exact here)."""
    from oracle.model import dice_from_counts
    g = torch.Generator().manual_seed(11)
    x = torch.rand(10, 4, 240, 240, generator=g)
    tg = (torch.rand(10, 1, 240, 240, generator=g) > 0.5).float()
    _, want = _oracle_counts(models, x, tg)
    merged = ysp.SegMetrics()
    per_slice = []
    for rank in range(2):
        lo, hi = ysp.shard_slices(2, 5, 2, rank)
        m = ysp.SegMetrics()
        for a, b in ysp.batches(lo, hi, 3):
            c = predictor.predict_raw(x[a:b].cuda(), tg[a:b].cuda())["counts"].cpu()
            per_slice.append(c)
            m.update(c)
        merged.state += m.state
    got = torch.cat(per_slice)
    d_got, d_want = ysp.dice_from_counts(got), dice_from_counts(want)
    assert (d_got - d_want).abs().max().item() <= 1e-4
    assert (got.long() - want).abs().max().item() <= 2            # a logit within 1e-5 of 0 may flip a pixel
    assert merged.compute()["slices"] == 10


def test_host_pipeline_matches_direct_call(ysp, predictor):
    g = torch.Generator().manual_seed(5)
    B = 6
    u8 = [torch.randint(0, 256, (B, 240, 240, 4), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    tg = (torch.rand(B, 1, 240, 240, generator=g) > 0.5).float().pin_memory()
    hp = ysp.HostPipeline(predictor, B, 240, 240)
    got = []
    for i in range(3):
        slot = hp.submit(u8[i], tg)
        got.append({k: v.clone() for k, v in hp.results(slot).items()})
    hp.synchronize()
    for i in range(3):
        o = predictor.predict_raw(u8[i].cuda(), tg.cuda())
        torch.cuda.synchronize()
        for k in ("counts", "det_count", "det_idx", "det_boxes"):
            n = o["det_count"].cpu()
            if k in ("det_idx", "det_boxes"):
                for b in range(B):
                    assert torch.equal(got[i][k][b, : n[b]], o[k].cpu()[b, : n[b]]), (i, k, b)
            else:
                assert torch.equal(got[i][k], o[k].cpu()), (i, k)
    assert hp.h2d_bytes == B * 240 * 240 * 4 + B * 240 * 240 * 4 and hp.d2h_bytes > 0


@pytest.mark.parametrize("mode", ["tc32", "bf16"])
def test_replicas_take_alternate_batches(ysp, models, mode):
    """Predictor(replicas=2): consecutive batches run on alternate engine handles / streams (submit_raw, HostPipeline) and
    every batch gets bit-identical results to the single-handle call, whatever overlaps with it."""
    pred, seg = models
    one = ysp.Predictor.from_modules(pred, seg, mode=mode)
    two = ysp.Predictor.from_modules(pred, seg, mode=mode, replicas=2)
    assert len(two.replicas) == 2 and two.replicas[0] is two
    g = torch.Generator().manual_seed(23)
    B, NB = 5, 6
    u8 = [torch.randint(0, 256, (B, 240, 240, 4), dtype=torch.uint8, generator=g) for _ in range(NB)]
    tg = (torch.rand(B, 240, 240, generator=g) > 0.5).to(torch.uint8) * 255
    keys = ("mask_logits", "counts", "det_count", "det_idx", "det_boxes", "y")
    want = []
    for x in u8:
        o = one.predict_raw(x.cuda(), tg.cuda())
        torch.cuda.synchronize()
        want.append({k: o[k].clone() for k in keys})
    # device-resident round robin
    d = [x.cuda() for x in u8]
    dt = tg.cuda()
    got = [None] * NB
    for i in range(NB):
        two.submit_raw(d[i], dt, after=lambda o, i=i: got.__setitem__(i, {k: o[k].clone() for k in keys}))
    two.join()
    torch.cuda.synchronize()
    assert two.replicas[1].engine.launches_total > 0 and two.launches_total == 2 * two.replicas[1].engine.launches_total
    for i in range(NB):
        n = want[i]["det_count"].tolist()
        for k in ("mask_logits", "counts", "det_count", "y"):
            assert torch.equal(got[i][k], want[i][k]), (i, k)
        for b in range(B):
            assert torch.equal(got[i]["det_idx"][b, : n[b]], want[i]["det_idx"][b, : n[b]])
            assert torch.equal(got[i]["det_boxes"][b, : n[b]], want[i]["det_boxes"][b, : n[b]])
    # host pipeline: slot k computes on replica k
    hp = ysp.HostPipeline(two, B, 240, 240)
    assert len(hp.run_streams) == 2
    hx, ht = [x.pin_memory() for x in u8], tg.pin_memory()
    res = []
    for i in range(NB):
        slot = hp.submit(hx[i], ht)
        res.append({k: v.clone() for k, v in hp.results(slot).items()})
    hp.synchronize()
    for i in range(NB):
        assert torch.equal(res[i]["counts"], want[i]["counts"].cpu()) and torch.equal(res[i]["det_count"], want[i]["det_count"].cpu())


@pytest.mark.parametrize("B,size", [(1, 240), (3, 160), (5, 96)])
def test_odd_batches_and_sizes(ysp, models, predictor, B, size):
    g = torch.Generator().manual_seed(B * 1000 + size)
    x = torch.rand(B, 4, size, size, generator=g)
    tg = (torch.rand(B, 1, size, size, generator=g) > 0.5).float()
    want_logits, want = _oracle_counts(models, x, tg)
    ml, dets, keep, counts = predictor.predict(x.cuda(), tg.cuda())
    assert (ml.cpu() - want_logits).abs().max().item() <= 1e-3
    assert (counts.cpu().long() - want).abs().max().item() <= 2
    assert len(dets) == B and all(d.shape[1] == 6 for d in dets)


def test_error_behaviour(ysp, models, predictor):
    pred, seg = models
    with pytest.raises(ysp.YspError):
        predictor.predict(torch.rand(1, 4, 240, 240))                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        predictor.predict(torch.rand(1, 3, 240, 240).cuda())                # 3 channels
    with pytest.raises(ValueError):
        predictor.predict(torch.rand(1, 4, 100, 100).cuda())                # H, W not multiples of 8
    eng = ysp.Engine("cuda:0", "fp32")
    sd = dict(seg.state_dict())
    sd.pop("decoder.3.1.conv.1.conv1.conv.weight")
    eng.load_state_dict("seg", sd)
    with pytest.raises(KeyError, match="decoder.3.1.conv.1.conv1"):
        eng.finalize(det=False, seg=True)


def test_objectmap_formats(ysp, models, tmp_path):
    """producer (generate_objectmaps.py:91-106) and consumer (dataset.py:86-97) of the bottleneck map."""
    from oracle.model import objectmap_transform as oref
    pred, _ = models
    det = ysp.B200Detector.from_predictor(pred, mode="fp32")
    x = torch.rand(3, 4, 160, 160, generator=torch.Generator().manual_seed(2))
    paths = ysp.save_objectmaps(det, x.cuda(), ["a", "b", "c"], str(tmp_path))
    with torch.no_grad():
        _, raws = pred.model(x)
    maps = []
    for i, p in enumerate(paths):
        m = torch.load(p)
        assert m.shape == (1, 1, 20, 20) and p.endswith("_20.pt")
        assert (m - raws[0][i:i + 1, -1:]).abs().max().item() <= 1e-3
        maps.append(m)
    maps = torch.cat(maps + [torch.full((1, 1, 20, 20), 0.75)])         # last map: std == 0 branch (exactly representable)
    got = ysp.objectmap_transform(maps.cuda()).cpu()
    want = torch.stack([oref(m) for m in maps])
    assert (got - want).abs().max().item() <= 1e-6


def test_scale_boxes(ysp):
    from oracle.model import scale_boxes as sref
    g = torch.Generator().manual_seed(4)
    b = torch.rand(37, 6, generator=g) * 300 - 20
    for img1, img0, padding in (((256, 256), (240, 240, 4), False), ((640, 640), (240, 320, 3), True), ((160, 160), (240, 240), True)):
        want = sref(img1, b[:, :4], img0, padding=padding)
        d = b.clone().cuda()
        out = ysp.scale_boxes(img1, d[:, :4], img0, padding=padding)
        assert torch.allclose(out.cpu(), want, atol=1e-5) and torch.equal(d[:, 4:].cpu(), b[:, 4:])


@pytest.mark.parametrize("mode", ["bf16", "tc32"])
def test_shared_stem_is_exact(ysp, models, monkeypatch, mode):
    """Tensor-core pipelines: reading the detector's layer-1 output through a pitched view (shared stem) gives bit-identical
    results to recomputing encoder layers 0-1 (YSP_NO_SHARE=1 at handle creation)."""
    pred, seg = models
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5, 4, 240, 240, generator=g).cuda()
    tg = (torch.rand(5, 1, 240, 240, generator=g) > 0.5).float().cuda()
    a = {k: v.clone() for k, v in ysp.Predictor.from_modules(pred, seg, mode=mode).predict_raw(x, tg).items()}
    monkeypatch.setenv("YSP_NO_SHARE", "1")
    P2 = ysp.Predictor.from_modules(pred, seg, mode=mode)
    monkeypatch.delenv("YSP_NO_SHARE")
    b = P2.predict_raw(x, tg)
    assert P2.engine.launches_total > 0
    assert torch.equal(a["mask_logits"], b["mask_logits"]) and torch.equal(a["counts"], b["counts"])
    assert torch.equal(a["det_idx"], b["det_idx"]) and torch.equal(a["y"], b["y"])


def test_confidence_gate_matches_sketch():
    """evaluate_model.py:149-155 (commented-out sketch): no detection, or best conf <= gate -> all-zero predicted mask."""
    import torch
    import yolo_u_b200 as ysp
    from oracle.model import build_models, synth_inputs
    pred, seg = build_models(0)
    P = ysp.Predictor.from_modules(pred, seg, device="cuda:0", mode="fp32")
    x, _, tg = synth_inputs(4, 240)
    x[1] = 0.0                                   # a blank slice: few / weak detections
    _, dets0, _, counts0 = P.predict(x.cuda(), tg.cuda())
    counts0 = counts0.clone()
    best = torch.tensor([float(d[0, 4]) if len(d) else -1.0 for d in dets0])
    gate = float(best.sort().values[1:3].mean())  # between the 2nd and 3rd best confidence: gates some, keeps some
    _, dets, _, counts = P.predict(x.cuda(), tg.cuda(), conf_gate=gate)
    want_gated = (best <= gate)
    assert want_gated.any() and not want_gated.all()
    assert torch.equal(P.gated.cpu().bool(), want_gated)
    for b in range(4):
        if want_gated[b]:
            assert counts[b, 0] == 0 and counts[b, 1] == 0 and counts[b, 2] == counts0[b, 2]
            assert int(P._out["mask"][b].sum()) == 0
        else:
            assert torch.equal(counts[b], counts0[b])
            assert int(P._out["mask"][b].sum()) == int(counts0[b, 1])
