"""Seg-head training step (BASELINE cfg 4) through the C ABI vs the oracle under torch autograd.

The oracle side is the PyTorch restatement of YOLOSegPlusPlus in train() mode (BN batch statistics) with the monai
DiceLoss restatement (oracle/model.py:dice_loss, train.py:98-104) and torch.optim.AdamW (train.py:262); the CUDA side is
libysp's hand-derived backward.  Tolerances (fp32, different summation orders): logits 1e-3 abs (north_star), loss
1e-5, gradients 2e-3 of the tensor's max-abs gradient, AdamW on given gradients 2e-6 abs."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B, S, seed=0, loss="dice"):
    from oracle.model import build_models, synth_inputs
    from yolo_u_b200.trainer import SegHeadTrainer
    _, seg = build_models(seed)
    x, lg, tg = synth_inputs(B, S, seed)
    tr = SegHeadTrainer(seg.state_dict(), batch_size=B, image_size=S, lr=1e-3, epochs=10, loss=loss, device="cuda:0")
    return seg, tr, x, lg, tg


def _oracle_loss(seg, x, lg, tg, kind):
    from oracle.model import dice_loss
    pred = seg(x, lg)
    loss = dice_loss(pred, tg)
    if kind == "dice_bce":
        loss = loss + torch.nn.functional.binary_cross_entropy_with_logits(pred, tg)
    return pred, loss


def _trainable(seg):
    return {k: p for k, p in seg.named_parameters() if not k.startswith("encoder.") and k != "param"}


@pytest.mark.parametrize("B,S,kind", [(2, 64, "dice"), (3, 240, "dice"), (2, 96, "dice_bce")])
def test_forward_backward_matches_autograd(B, S, kind):
    seg, tr, x, lg, tg = _setup(B, S, loss=kind)
    seg.train()
    for p in _trainable(seg).values():
        p.requires_grad_(True)
    pred_ref, loss_ref = _oracle_loss(seg, x, lg, tg, kind)
    loss_ref.backward()
    loss3, pred = tr.forward_backward(x.cuda(), tg.cuda(), lg.cuda())
    torch.cuda.synchronize()
    assert (pred.cpu() - pred_ref.detach()).abs().max().item() <= 1e-3
    assert abs(loss3[0].item() - loss_ref.item()) <= 1e-5
    grads = tr.named_grads()
    ref = _trainable(seg)
    assert set(grads) == set(ref)
    worst = 0.0
    # some gradients are exactly zero in exact arithmetic (a bias in front of a train-mode BN): floor the tolerance
    # at 1e-5 of the largest gradient of the model
    floor = 1e-5 * max(p.grad.abs().max().item() for p in ref.values())
    for k, p in ref.items():
        g_ref, g = p.grad, grads[k].cpu()
        scale = g_ref.abs().max().item()
        err = (g - g_ref).abs().max().item()
        worst = max(worst, err / (scale + 1e-12))
        assert err <= 2e-3 * scale + floor, f"{k}: err {err:.3e} vs scale {scale:.3e}"
    # BN running statistics and num_batches_tracked follow nn.BatchNorm2d.train()
    sd, sd_ref = tr.state_dict(), seg.state_dict()
    for k, v in sd_ref.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert torch.allclose(sd[k].cpu(), v, rtol=1e-4, atol=1e-5), k
        if k.endswith("num_batches_tracked") and k.startswith("decoder."):
            assert int(sd[k]) == int(v), k


def _adamw(p, g, m, v, step, lr=1e-3, eps=1e-8, wd=1e-2, gscale=1.0, max_norm=0.0):
    from yolo_u_b200._lib import check, lib
    scratch = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(lib().ysp_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, 0.9, 0.999, eps, wd, step,
                          gscale, max_norm, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream))


@pytest.mark.parametrize("max_norm,gscale,eps", [(0.0, 1.0, 1e-8), (0.0, 0.5, 1e-3), (0.05, 1.0, 1e-3)])
def test_adamw_kernel_matches_torch_optimizer(max_norm, gscale, eps):
    """ysp_adamw on given gradients vs torch.optim.AdamW (+ clip_grad_norm_ when max_norm > 0; eps 1e-3 makes the
    update sensitive to the gradient scale, which plain Adam is not)."""
    g0 = torch.Generator().manual_seed(5)
    n = 63764
    p_ref = torch.nn.Parameter(torch.randn(n, generator=g0))
    opt = torch.optim.AdamW([p_ref], lr=1e-3, eps=eps)
    p = p_ref.detach().clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(n, generator=g0) * 10 ** torch.empty(n).uniform_(-6, 0, generator=g0)
        p_ref.grad = (g * gscale).clone()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_([p_ref], max_norm=max_norm)
            assert total > max_norm
        opt.step()
        _adamw(p, g.cuda(), m, v, step, eps=eps, gscale=gscale, max_norm=max_norm)
    assert (p.cpu() - p_ref.detach()).abs().max().item() <= 2e-6


def test_training_steps_track_torch_optimizer():
    """3 full iterations (forward, loss, backward, AdamW) next to the oracle + torch.optim.AdamW.  Adam normalises every
    gradient by its own magnitude, so elements whose gradient is at rounding-noise level move by +-lr on either side:
    the trajectories are compared through the losses and the relative L2 distance of the parameters."""
    B, S = 2, 64
    seg, tr, x, lg, tg = _setup(B, S)
    seg.train()
    params = [p for p in _trainable(seg).values()]
    for p in params:
        p.requires_grad_(True)
    start = {k: p.detach().clone() for k, p in _trainable(seg).items()}
    opt = torch.optim.AdamW(params, lr=1e-3)
    for _ in range(3):
        opt.zero_grad()
        _, loss = _oracle_loss(seg, x, lg, tg, "dice")
        loss.backward()
        opt.step()
        l3, _ = tr.step(x.cuda(), tg.cuda(), lg.cuda())
        assert abs(l3[0].item() - loss.item()) <= 2e-5
    mine = tr.named_parameters()
    num = den = 0.0
    for k, p in _trainable(seg).items():
        num += (mine[k].cpu() - p.detach()).pow(2).sum().item()
        den += (p.detach() - start[k]).pow(2).sum().item()
    assert den > 0 and (num / den) ** 0.5 <= 2e-2, (num, den)


def test_state_dict_roundtrip_and_eval_engine():
    """`best.pth` compatibility (train.py:428 -> evaluate_model.py:243): the trainer's state_dict loads strictly into
    the reference module restatement, and the inference engine over it reproduces the oracle in eval mode."""
    B, S = 2, 64
    seg, tr, x, lg, tg = _setup(B, S)
    for _ in range(2):
        tr.step(x.cuda(), tg.cuda(), lg.cuda())
    sd = tr.state_dict()
    assert set(sd) == set(seg.state_dict())
    seg2 = copy.deepcopy(seg)
    seg2.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=True)
    seg2.eval()
    with torch.no_grad():
        ref = seg2(x, lg)
    out = tr.eval_engine("fp32").segpp_forward(x.cuda(), lg.cuda())
    assert (out.cpu() - ref).abs().max().item() <= 1e-3


def test_training_reduces_loss_and_scheduler():
    import math
    B, S = 4, 64
    seg, tr, x, lg, tg = _setup(B, S)
    blob = torch.zeros_like(tg)
    blob[:, :, 16:48, 20:44] = 1.0
    first = last = None
    for i in range(40):
        l3, _ = tr.step(x.cuda(), blob.cuda(), lg.cuda())
        if i == 0:
            first = l3[0].item()
    last = l3[0].item()
    assert last < first - 0.05, (first, last)
    lr = tr.scheduler_step()
    assert abs(lr - 0.5 * 1e-3 * (1 + math.cos(math.pi / 10))) < 1e-12
    assert tr.launches_per_step > 100          # the step is made of this library's kernels


def test_shape_errors():
    seg, tr, x, lg, tg = _setup(2, 64)
    with pytest.raises(ValueError):
        tr.forward_backward(x[:1].cuda(), tg.cuda(), lg.cuda())
    with pytest.raises(RuntimeError):
        tr.forward_backward(x.cuda(), tg.cuda(), lg[:, :, :4].cuda())
    with pytest.raises(Exception):
        tr.forward_backward(x, tg, lg)            # CPU tensors: no fallback


def test_validate_batch_matches_eval_mode_oracle():
    """train.py:346-366: eval-mode forward + loss value + Dice counters after a few training steps."""
    import copy
    from oracle.model import dice_loss, mask_counts
    B, S = 2, 64
    seg, tr, x, lg, tg = _setup(B, S)
    for _ in range(2):
        tr.step(x.cuda(), tg.cuda(), lg.cuda())
    loss3, counts, pred = tr.validate_batch(x.cuda(), tg.cuda(), lg.cuda())
    seg2 = copy.deepcopy(seg)
    seg2.load_state_dict({k: v.cpu() for k, v in tr.state_dict().items()}, strict=True)
    seg2.eval()
    with torch.no_grad():
        ref = seg2(x, lg)
        want = dice_loss(ref, tg).item()
    assert (pred.cpu() - ref).abs().max().item() <= 1e-3
    assert abs(loss3[0].item() - want) <= 1e-5
    assert torch.equal(counts.cpu().long(), mask_counts(pred.cpu(), tg))


def test_mixed_precision_branch_grad_scaler():
    """train.py:302-341 (the branch the reference runs by default): scaler.scale(loss).backward(), unscale_, scaler.step that
    SKIPS the optimiser when a gradient is inf/NaN, scaler.update (backoff 0.5 after a skip, x2 after growth_interval clean
    steps).  Tensors stay fp32 here, so a clean scaled step must equal the plain step up to the summation order of the
    gradient atomics (power-of-two scaling is exact)."""
    from yolo_u_b200.trainer import SegHeadTrainer
    B, S = 2, 64
    seg, plain, x, lg, tg = _setup(B, S)
    amp = SegHeadTrainer(seg.state_dict(), batch_size=B, image_size=S, lr=1e-3, epochs=10, device="cuda:0",
                         mixed_precision=True, growth_interval=2)
    assert amp.scale == 2.0 ** 16 and plain.scale == 1.0
    xc, lgc, tgc = x.cuda(), lg.cuda(), tg.cuda()
    # scaled backward: gradients are 2**16 x the plain ones, the loss value is not scaled
    l_amp, _ = amp.forward_backward(xc, tgc, lgc, grad_scale=amp.scale)
    l_plain, _ = plain.forward_backward(xc, tgc, lgc)
    assert abs(l_amp[0].item() - l_plain[0].item()) <= 1e-6
    g_amp, g_plain = amp.grads / amp.scale, plain.grads
    assert (g_amp - g_plain).abs().max().item() <= 2e-4 * g_plain.abs().max().item()
    # clean steps: same trajectory as the plain trainer; the scale doubles after growth_interval = 2 of them
    assert amp.optimizer_step() is True and plain.optimizer_step() is True
    amp.step(xc, tgc, lgc)
    plain.step(xc, tgc, lgc)
    assert amp.scale == 2.0 ** 17 and amp.step_count == 2 and amp.skipped_steps == 0
    start = SegHeadTrainer(seg.state_dict(), batch_size=B, image_size=S, device="cuda:0").params
    rel = ((amp.params - plain.params).norm() / (plain.params - start).norm()).item()
    assert rel <= 2e-2, rel            # Adam turns rounding-level gradient differences into +-lr moves (see above)
    # an overflowing gradient: the step is skipped (parameters, moments and step count untouched), the scale backs off
    before, m_before = amp.params.clone(), amp.adam_m.clone()
    amp.forward_backward(xc, tgc, lgc, grad_scale=amp.scale)
    amp.grads[7] = float("inf")
    assert amp.optimizer_step() is False
    assert torch.equal(amp.params, before) and torch.equal(amp.adam_m, m_before)
    assert amp.scale == 2.0 ** 16 and amp.step_count == 2 and amp.skipped_steps == 1
    amp.forward_backward(xc, tgc, lgc, grad_scale=amp.scale)
    amp.grads[3] = float("nan")
    assert amp.optimizer_step() is False and amp.scale == 2.0 ** 15
    # and training goes on
    l3, _ = amp.step(xc, tgc, lgc)
    assert amp.step_count == 3 and torch.isfinite(l3).all()


def test_encoder_lookahead_gives_identical_steps():
    """step(..., next_img=...) launches the NEXT batch's frozen-encoder pass on a side stream; the training trajectory does
    not change (same losses; parameters equal up to the summation order of the gradient atomics)."""
    B, S = 2, 64
    seg, a, x, lg, tg = _setup(B, S)
    _, b, _, _, _ = _setup(B, S)
    from oracle.model import synth_inputs
    batches = [tuple(t.cuda() for t in synth_inputs(B, S, seed=40 + i)) for i in range(4)]
    la, lb = [], []
    for i, (xi, lgi, tgi) in enumerate(batches):
        la.append(a.step(xi, tgi, lgi)[0].clone())
        nxt = batches[i + 1][0] if i + 1 < len(batches) else None
        lb.append(b.step(xi, tgi, lgi, next_img=nxt)[0].clone())
    torch.cuda.synchronize()
    assert b._ahead is None and b._enc_stream is not None
    for u, v in zip(la, lb):
        assert (u - v).abs().max().item() <= 2e-5
    start = _setup(B, S)[1].params
    rel = ((a.params - b.params).norm() / (a.params - start).norm()).item()
    assert rel <= 2e-2, rel
    # a batch that was not announced still works (encoder runs inline), and an announced batch that never arrives is dropped
    b.step(batches[0][0], batches[0][2], batches[0][1], next_img=batches[1][0])
    l = b.step(batches[2][0], batches[2][2], batches[2][1])[0]
    torch.cuda.synchronize()
    assert torch.isfinite(l).all()
