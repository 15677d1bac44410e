"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference via oracle/refshim.py) on seeded synthetic inputs.

Run in the build container:   python -m oracle.make_golden
The GPU box never runs this (no /root/reference there); it only reads the committed vectors.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from lgd_b200 import synth
from oracle import refshim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (cfg kwargs, batch kwargs, distill_flag)
    "ctx_stu_adv": (dict(add_context_box=True, interact_pattern="stuGuided"),
                    dict(B=2, img_h=120, img_w=150, seed=101, adversarial=True), 1),
    "noctx_stu_empty": (dict(add_context_box=False, interact_pattern="stuGuided"),
                        dict(B=3, img_h=100, img_w=130, seed=102, n_boxes=[4, 0, 9]), 1),
    "ctx_label_detach": (dict(add_context_box=True, interact_pattern="labelGuided", detach_appearance_embed=True),
                         dict(B=2, img_h=128, img_w=190, seed=103, n_boxes=[0, 6]), 0),
    # BOX_FORMAT x1y1wh (utils.py:26-38): the synthetic XYXY tensors are read as (x1, y1, w, h)
    "ctx_stu_x1y1wh": (dict(add_context_box=True, interact_pattern="stuGuided", box_format="x1y1wh"),
                       dict(B=2, img_h=128, img_w=160, seed=104, n_boxes=[0, 5]), 1),
    # the Mask R-CNN recipe (configs/Distillation/MaskRCNN: LOAD_LABELMAP True, DETACH_APPEARANCE_EMBED True): 133-dim
    # descriptors, polygon-mask pooling / rendering; one image smaller than the padded batch, one without GT
    "seg_ctx_detach": (dict(add_context_box=True, interact_pattern="stuGuided", detach_appearance_embed=True,
                            load_labelmap=True),
                       dict(B=3, img_h=128, img_w=160, seed=105, n_boxes=[4, 0, 6], with_masks=True,
                            unpadded=[(128, 160), (120, 150), (100, 160)]), 1),
    # CATEGORY_FORMAT norm_classes (label_encoder.py:24-25,91-93): 5-dim descriptors (box + class index / 80). Only
    # without the context box: with it the reference concatenates (N+1, 4) boxes with (N, 1) classes and raises.
    "noctx_stu_normcls": (dict(add_context_box=False, interact_pattern="stuGuided", category_format="norm_classes"),
                          dict(B=3, img_h=100, img_w=130, seed=106, n_boxes=[5, 0, 3]), 1),
}
WEIGHT_SEED = 5


def run_case(name):
    cfg_kw, batch_kw, flag = CASES[name]
    cfg = synth.make_cfg(**cfg_kw)
    R = refshim.RefDistillator(cfg)
    sd = synth.synth_state_dict(WEIGHT_SEED, desc_dim=synth.desc_dim_of(cfg_kw))
    missing = R.teacher.load_state_dict({k[len("teacher."):]: v for k, v in sd.items() if k.startswith("teacher.")})
    R.D.adapter.load_state_dict({k[len("adapter."):]: v for k, v in sd.items() if k.startswith("adapter.")})
    bi, im, feats = synth.synth_batch(requires_grad=True, **batch_kw)
    cap = {"mha_q": [], "mha_kv": [], "mha_out": []}
    def _h_le(m, i, o):
        cap["label_embed"] = o[0].detach()

    def _h_cp(m, i, o):
        cap["canoni"] = o.detach()

    def _h_mha(m, i, o):
        cap["mha_q"].append(i[0].detach().squeeze(1))
        cap["mha_kv"].append(i[1].detach().squeeze(1))
        cap["mha_out"].append(o[0].detach().squeeze(1))

    hooks = [R.teacher.label_encoder_.register_forward_hook(_h_le),
             R.teacher.canoni_proj_1D.register_forward_hook(_h_cp),
             R.teacher.multi_head_attn.register_forward_hook(_h_mha)]
    tea, inst_labels, masks, loss = R.step(bi, im, feats, distill_flag=flag)
    for h in hooks:
        h.remove()
    cot = synth.synth_cotangents(tea)
    total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
    named = [("teacher." + k, p) for k, p in R.teacher.named_parameters()] + \
            [("adapter." + k, p) for k, p in R.D.adapter.named_parameters()]
    grads = torch.autograd.grad(total, list(feats.values()) + [p for _, p in named], allow_unused=True)
    gfeat = grads[:len(feats)]
    gparam = grads[len(feats):]
    out = {"loss": loss.detach().numpy(), "weight_seed": np.int64(WEIGHT_SEED), "distill_flag": np.int64(flag)}
    out["label_embed"] = cap["label_embed"].numpy()
    out["canoni"] = cap["canoni"].numpy()
    for l, k in enumerate(feats):
        out[f"feat_sum_{k}"] = feats[k].detach().double().sum().numpy()
        out[f"tea_{k}"] = tea[k].detach().numpy()
        out[f"gfeat_{k}"] = (gfeat[l].numpy() if gfeat[l] is not None else np.zeros(0, np.float32))
        m = torch.cat(masks[l], 0)
        out[f"mask_{k}"] = np.packbits(m.numpy().astype(np.uint8), axis=1)
        out[f"mask_shape_{k}"] = np.array(m.shape)
        if cap["mha_q"]:
            out[f"mha_q_{k}"] = cap["mha_q"][l].numpy()
            out[f"mha_kv_{k}"] = cap["mha_kv"][l].numpy()
            out[f"mha_out_{k}"] = cap["mha_out"][l].numpy()
    out["counts"] = np.array([m.shape[0] for m in masks[0]])
    for i, il in enumerate(inst_labels):
        out[f"inst_labels_{i}"] = il.numpy().astype(np.float32)
    # parameter gradients: full tensors for the small ones, norms + strided samples for all
    for (n, p), g in zip(named, gparam):
        if g is None:
            out["gnone_" + n] = np.int64(1)
            continue
        out["gnorm_" + n] = g.double().norm().numpy()
        flat = g.reshape(-1)
        stride = max(1, flat.numel() // 4096)
        out["gsamp_" + n] = flat[::stride].numpy()
    out["wsum"] = np.array([float(v.double().sum()) for _, v in sorted(sd.items())])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss), "T", int(out["counts"].sum()), "saved")


# the student's FCOS-family head on a (teacher) pyramid, SURVEY.md 8(f) rank 1: the reference's own FCOSHead / POTOHead
# classes (thirdparty_heads/fcos.py:433-546, poto.py:523-625) on seeded inputs
HEAD_CASES = {
    # name: (class, centerness_on_reg, norm_reg_targets, B, level sizes, strides, weight seed, input seed)
    "fcos_head_ctr_on_reg": ("FCOSHead", True, True, 2, [(12, 16), (6, 8), (3, 4)], [8, 16, 32], 21, 31),
    "fcos_head_ctr_on_cls_exp": ("FCOSHead", False, False, 1, [(10, 14), (5, 7)], [8, 16], 22, 32),
    "poto_head": ("POTOHead", True, True, 2, [(9, 12), (5, 6)], [8, 16], 23, 33),
}


def head_inputs(name):
    cls, ctr_on_reg, norm_reg, B, hws, strides, wseed, xseed = HEAD_CASES[name]
    sd = synth.synth_fcos_head_state_dict(wseed, len(hws), 80, centerness=(cls == "FCOSHead"))
    gen = torch.Generator().manual_seed(xseed)
    feats = [torch.randn(B, 256, h, w, generator=gen) for (h, w) in hws]
    # cotangents of the three outputs (the losses' gradients): (logits, bbox_reg, centerness) per level
    cots = [[torch.randn(B, c, h, w, generator=gen) * 1e-3 for (h, w) in hws] for c in (80, 4, 1)]
    return sd, feats, cots


def run_head_case(name):
    import types
    cls, ctr_on_reg, norm_reg, B, hws, strides, wseed, xseed = HEAD_CASES[name]
    H = refshim.load_heads()
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(FCOS=types.SimpleNamespace(
        NUM_CLASSES=80, NUM_CONVS=4, PRIOR_PROB=0.01, FPN_STRIDES=strides, CENTERNESS_ON_REG=ctr_on_reg,
        NORM_REG_TARGETS=norm_reg)))
    head = getattr(H, cls)(cfg, [H.ShapeSpec(channels=256)] * len(hws))
    sd, feats, cots = head_inputs(name)
    head.load_state_dict(sd)
    fx = [f.clone().requires_grad_(True) for f in feats]
    outs = head(fx)
    total = sum((o * c).sum() for group, cg in zip(outs, cots) for o, c in zip(group, cg))
    named = list(head.named_parameters())
    grads = torch.autograd.grad(total, fx + [p for _, p in named])
    out = {"wsum": np.array([float(v.double().sum()) for _, v in sorted(sd.items())]),
           "feat_sum": np.array([float(f.double().sum()) for f in feats])}
    for gi, gname in enumerate(("logits", "bbox_reg", "centerness")[:len(outs)]):
        for l, o in enumerate(outs[gi]):
            out["%s_%d" % (gname, l)] = o.detach().numpy()
    for l in range(len(hws)):
        out["gfeat_%d" % l] = grads[l].numpy()
    for (n, p), g in zip(named, grads[len(fx):]):
        out["gnorm_" + n] = g.double().norm().numpy()
        flat = g.reshape(-1)
        out["gsamp_" + n] = flat[::max(1, flat.numel() // 4096)].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "saved")


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    for n in CASES:
        run_case(n)
    for n in HEAD_CASES:
        run_head_case(n)
