"""Generate tests/golden/segpp_golden.pt by running the REFERENCE's own files, unmodified, in this container:

  * /root/reference/YOLOSegPlusPlus.py   (DoubleLightConv :33-58, ECA :60-88, decoder topology :150-178, forward :242-272)
  * /root/reference/_YOLOSegPlusPlus.py  (the no-logit ablation, :157 and :264-268)
  * /root/reference/dataset.py:89-93,97  (objectmap z-score + sigmoid, executed verbatim as a source fragment)
  * /root/reference/evaluate_model.py:157-158,166-168,177-178  (mask threshold, TP / FP / FN, precision / recall)

through oracle/ref_shim.py (which stubs the two absent imports; the upstream ultralytics blocks behind
`ultralytics.nn.modules` are the restatement in oracle/modules.py -- un-vendored, so they stay unpinned).

Run here (the only place /root/reference exists):   python tests/golden/make_segpp_golden.py
The fixture stores seeds / recipes and the reference OUTPUTS; tests rebuild weights and inputs from the seeds
(oracle.model.build_models / synth_init_ / synth_inputs are deterministic torch-CPU generators).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

REF = "/root/reference"
SHAPES = ((1, 240, 240), (1, 160, 160), (2, 64, 96))          # BASELINE size, the reference's native 160, a non-square %8
ABL_SEED = 11


def seg_inputs(b, h, w):
    g = torch.Generator().manual_seed(1000 + h + w)
    x = torch.rand(b, 4, h, w, generator=g)
    lg = torch.sigmoid(torch.randn(b, 1, h // 8, w // 8, generator=g))
    return x, lg


def objectmaps():
    g = torch.Generator().manual_seed(321)
    maps = [torch.randn(1, 1, 20, 20, generator=g) * 3 - 4, torch.randn(1, 1, 30, 30, generator=g) * 0.1, torch.full((1, 1, 20, 20), -2.5)]
    return maps                                               # the last one has std == 0 (dataset.py:93)


def metric_case():
    g = torch.Generator().manual_seed(654)
    pred = torch.randn(3, 1, 64, 64, generator=g)
    pred[0, 0, 0, :4] = torch.tensor([0.0, 5e-8, -5e-8, 1e-6])                # the sigmoid == 0.5 band
    mask = (torch.rand(3, 1, 64, 64, generator=g) > 0.5).float()
    mask[2] = 0
    return pred, mask


def main():
    from oracle.model import build_models, synth_init_
    from oracle.ref_shim import load_reference_segpp, reference_lines, segpp_ctor_patch
    out = {"shapes": SHAPES, "abl_seed": ABL_SEED}
    pred, oseg = build_models(0)

    # ---- YOLOSegPlusPlus.py ------------------------------------------------------------------------------------------
    ref = load_reference_segpp(os.path.join(REF, "YOLOSegPlusPlus.py"))
    with segpp_ctor_patch():
        m = ref.YOLOSegPlusPlus(pred)
    missing, unexpected = m.load_state_dict(oseg.state_dict(), strict=True)
    m.eval()
    out["state_dict_keys"] = sorted(m.state_dict().keys())
    out["head_params"] = sum(p.numel() for n, p in m.named_parameters() if not n.startswith("encoder."))
    outs = []
    with torch.no_grad():
        for (b, h, w) in SHAPES:
            x, lg = seg_inputs(b, h, w)
            outs.append(m(x, lg).clone())
        # the decoder stages of the 240 case: pins concat order / skip pops stage by stage
        x, lg = seg_inputs(*SHAPES[0])
        stages = {}
        hooks = [m.decoder[i].register_forward_hook(lambda mod, i_, o, k=i: stages.__setitem__(k, o.clone())) for i in range(5)]
        m(x, lg)
        for hk in hooks:
            hk.remove()
    out["logits"] = outs
    out["stage_stats"] = {k: (float(v.double().mean()), float(v.double().std()), float(v.abs().max())) for k, v in stages.items()}
    out["stage_samples"] = {k: v.flatten()[:: max(v.numel() // 4096, 1)][:4096].clone() for k, v in stages.items()}

    # ---- _YOLOSegPlusPlus.py (ablation: decoder.0 = C3Ghost(128, 96), input = skip only) --------------------------------
    abl = load_reference_segpp(os.path.join(REF, "_YOLOSegPlusPlus.py"))
    with segpp_ctor_patch():
        ma = abl.YOLOSegPlusPlus(pred)
    sd = {k: v for k, v in oseg.state_dict().items() if not k.startswith("decoder.0.0.")}
    ma.load_state_dict(sd, strict=False)
    synth_init_(ma.decoder[0][0], ABL_SEED, lin_gain=2.0)
    ma.eval()
    with torch.no_grad():
        x, lg = seg_inputs(2, 96, 96)
        out["abl_logits"] = ma(x, lg).clone()

    # ---- dataset.py:89-93 + :97, verbatim ---------------------------------------------------------------------------------
    frag = reference_lines(os.path.join(REF, "dataset.py"), 89, 93, "objectmap_tensor.std()")
    ret = reference_lines(os.path.join(REF, "dataset.py"), 97, 97, "torch.sigmoid(objectmap_tensor)")
    assert ret.strip().startswith("return img_tensor, mask_tensor, torch.sigmoid(objectmap_tensor)")
    res = []
    for mp in objectmaps():
        ns = {"torch": torch, "objectmap_tensor": mp.squeeze(0)}          # :86  torch.load(path).squeeze(0)
        exec(frag, ns)
        res.append(torch.sigmoid(ns["objectmap_tensor"]).clone())         # :97
    out["objectmap_out"] = res

    # ---- evaluate_model.py:157-158, 166-168, 177-178, verbatim -----------------------------------------------------------
    ev = os.path.join(REF, "evaluate_model.py")
    f1 = reference_lines(ev, 157, 158, "pred_binary  = (pred_sigmoid > 0.5).float()")
    f2 = reference_lines(ev, 166, 168, "FN = ((1 - pred_binary) * mask).sum().float()")
    f3 = reference_lines(ev, 177, 178, "val_recall_metric")
    pred_l, mask = metric_case()
    per = []
    tot = [0.0, 0.0, 0.0]
    for i in range(pred_l.shape[0]):                                      # the reference evaluates with batch size 1 (:255)
        ns = {"torch": torch, "pred": pred_l[i:i + 1], "mask": mask[i:i + 1]}
        exec(f1, ns)
        exec(f2, ns)
        per.append((ns["TP"].item(), ns["FP"].item(), ns["FN"].item()))
        for j, k in enumerate(("TP", "FP", "FN")):
            tot[j] += ns[k].item()
        if i == 0:
            out["pred_binary0"] = ns["pred_binary"].clone()
    ns = {"total_TP": tot[0], "total_FP": tot[1], "total_FN": tot[2]}
    exec(f3, ns)
    out["tp_fp_fn"] = per
    out["precision_recall"] = (ns["val_precision_metric"], ns["val_recall_metric"])

    dst = os.path.join(ROOT, "tests", "golden", "segpp_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes;", "head params", out["head_params"], "missing", missing, "unexpected", unexpected)


if __name__ == "__main__":
    main()
