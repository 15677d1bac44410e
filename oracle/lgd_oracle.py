"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of LGD's distillation hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product path (lgd_b200/) never does and fails loudly without its CUDA
library.

Parity status: PINNED AGAINST THE REFERENCE ITSELF, not against reference-owned tests (the
reference ships none). oracle/make_golden.py imports the unmodified
reference through oracle/refshim.py in the build container, runs it on seeded inputs and commits
the outputs under tests/golden/; tests/test_oracle.py checks this restatement against those
vectors (masks bit-exact, floats to fp32 round-off).

It is a staged, functional restatement (plain torch CPU ops, fp32 or fp64) of:
  a1  box_descriptor_encode      models/customized_detectors/dynamic_teacher/label_encoder.py:12-115
  a2  LabelEncoder.forward/STN   label_encoder.py:216-276, spatial_transformer.py:30-47
  a3  canoni_proj_1D, student_proj_2D   dynamic_teacher.py:229-235, layers.py:9-32
  a4  get_inside_gt_mask         dynamic_teacher/utils.py:53-89
  a5  aggregate_per_level        dynamic_teacher.py:81-103
  a6  MultiheadAttention block   dynamic_teacher.py:255-275
  a7  rendering                  dynamic_teacher.py:106-206
  a8  refinement_module          dynamic_teacher.py:67-73,280-281
  a10 SequentialConvs            models/adapters/sequential_convs.py:7-15
  a11 BaseDistillator.distill    models/base_distillator.py:34-64
Weights are taken from a state_dict with the reference's checkpoint names.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

EPS = 1e-5
NUM_CLASSES = 80


# ----------------------------------------------------------------------------- a1: box table
def prepare_boxes(instances: Sequence, img_h: int, img_w: int, add_context_box: bool, box_format: str = "x1y1x2y2",
                  add_mask: bool = False, category_format: str = "one_hot"):
    """Per image: (boxes (N_i,4) fp32 clamped, onehot (N_i,80) fp32, inst_labels[, mask descriptors (N_i,49)]).
    add_mask (LOAD_LABELMAP, label_encoder.py:31-32,60-69,79-80): 7x7 box-relative bitmask of every instance from
    gt_masks.crop_and_resize, all ones for the context row, all zeros for the dummy row of an image without GT.

    label_encoder.py:40-99. Zero-GT image -> one dummy box [0,0,1,1], zero one-hot, float label
    [0.] and NO context box (:57-69,:75). Context box [0,0,W,H] appended before clamping (:75-83);
    its one-hot row is all zero because scatter_ uses the un-extended labels (:99).
    category_format 'norm_classes' (:24-25,91-93): the class slot is ONE column, class index / num_classes; with the
    context box the reference concatenates (N+1, 4) boxes with (N, 1) classes and torch.cat raises (:105) -- so does this.
    """
    if category_format not in ("one_hot", "norm_classes"):
        raise ValueError('Unsupported class_descriptor mode: {} !'.format(category_format))
    out = []
    for inst in instances:
        n = len(inst)
        if n > 0:
            b = inst.gt_boxes.tensor.reshape(n, 4).to(torch.float32).cpu().clone()
            labels = inst.gt_classes.reshape(n).cpu()
            assert bool(((labels >= 0) & (labels <= NUM_CLASSES - 1)).all()), "label out of range"
            m49 = None
            if add_mask:
                m49 = inst.gt_masks.crop_and_resize(inst.gt_boxes.tensor, 7).reshape(n, 49).to(torch.float32)
            if box_format == "x1y1wh":      # utils.py:26-38, before the context box (label_encoder.py:72-77)
                b = torch.stack([b[:, 0], b[:, 1], b[:, 0] + b[:, 2] - 1.0, b[:, 1] + b[:, 3] - 1.0], 1)
            if add_context_box:
                b = torch.cat([b, torch.tensor([[0.0, 0.0, float(img_w), float(img_h)]])], 0)
                if add_mask:
                    m49 = torch.cat([m49, torch.ones(1, 49)], 0)
            if category_format == "norm_classes":
                if add_context_box:
                    raise RuntimeError("Sizes of tensors must match except in dimension 1. Expected size %d but got size "
                                       "%d for tensor number 1 in the list." % (n + 1, n))
                onehot = labels.reshape(n, 1) / NUM_CLASSES      # int64 / int -> fp32 true division
            else:
                onehot = torch.zeros(b.shape[0], NUM_CLASSES)
                onehot[torch.arange(n), labels] = 1.0
            inst_labels = labels
        else:
            b = torch.tensor([[0.0, 0.0, 1.0, 1.0]])
            if box_format == "x1y1wh":
                b = torch.stack([b[:, 0], b[:, 1], b[:, 0] + b[:, 2] - 1.0, b[:, 1] + b[:, 3] - 1.0], 1)
            onehot = torch.zeros(1, 1 if category_format == "norm_classes" else NUM_CLASSES)   # zeros / 80 = 0
            inst_labels = torch.zeros(1)
            m49 = torch.zeros(1, 49) if add_mask else None
        # clamp_x1y1x2y2, utils.py:40-51
        b = torch.stack([b[:, 0].clamp(0, img_w - 1), b[:, 1].clamp(0, img_h - 1),
                         b[:, 2].clamp(0, img_w - 1), b[:, 3].clamp(0, img_h - 1)], 1)
        out.append((b, onehot, inst_labels, m49) if add_mask else (b, onehot, inst_labels))
    return out


def encode_descriptors(boxes: torch.Tensor, onehot: torch.Tensor, img_h: int, img_w: int, m49=None):
    """(N,4)+(N,80)[+(N,49)] -> (N,84 | 133) in [-1,1]. label_encoder.py:88-112, utils.py:16-24 (fp32)."""
    nb = boxes.clone()
    nb[:, [0, 2]] = nb[:, [0, 2]] / float(img_w)
    nb[:, [1, 3]] = nb[:, [1, 3]] / float(img_h)
    d = torch.cat([nb, onehot], 1)
    if m49 is not None:
        d = torch.cat([d, m49], 1)
    assert bool(((d >= 0) & (d <= 1)).all()), "descriptor outside [0,1]"
    return 2.0 * (d - 0.0) + (-1.0)


# ----------------------------------------------------------------------------- helpers
def _ln(x):  # LayerNorm over the last dim, no affine, biased var, eps 1e-5
    return F.layer_norm(x, (x.shape[-1],), eps=EPS)


def _lin(x, sd, name):
    w = sd[name + ".weight"]
    if w.dim() == 3:  # Conv1d(k=1) on a length-1 sequence == Linear (label_encoder.py:149-155)
        w = w[:, :, 0]
    return F.linear(x, w.to(x.dtype), sd[name + ".bias"].to(x.dtype))


def _stn(x, sd, p, k):
    """spatial_transformer.py:30-47 on (T,k) rows -> (T,k,k). No identity shortcut (:42-44)."""
    h = F.relu(_ln(_lin(x, sd, p + ".conv1")))
    h = F.relu(_ln(_lin(h, sd, p + ".conv2")))
    h = F.relu(_ln(_lin(h, sd, p + ".conv3")))
    h = F.relu(_ln(_lin(h, sd, p + ".fc1")))
    h = F.relu(_ln(_lin(h, sd, p + ".fc2")))
    h = _lin(h, sd, p + ".fc3")
    return h.view(-1, k, k)


def label_encoder(desc: torch.Tensor, counts: Sequence[int], sd: Dict[str, torch.Tensor],
                  prefix: str = "teacher.label_encoder_"):
    """a2. desc (T,84) -> label embeddings (T,256). label_encoder.py:239-274 with R=1."""
    p = prefix
    t_desc = _stn(desc, sd, p + ".stn_desc", desc.shape[1])
    x = torch.bmm(desc.unsqueeze(1), t_desc).squeeze(1)                    # :241
    x = F.relu(_ln(_lin(x, sd, p + ".conv1")))                             # :243
    t_feat = _stn(x, sd, p + ".stn_feat", 64)
    x_ft = torch.bmm(x.unsqueeze(1), t_feat).squeeze(1)                    # :248
    x = F.relu(_ln(_lin(x_ft, sd, p + ".conv2")))
    x = F.relu(_ln(_lin(x, sd, p + ".conv3")))                             # (T,1024)
    pooled = torch.stack([c.max(dim=0)[0] for c in x.split(list(counts), 0)], 0)   # :195-213
    xg = torch.cat([pooled[i:i + 1].expand(n, -1) for i, n in enumerate(counts)], 0)
    x = F.relu(_ln(_lin(torch.cat([x_ft, xg], 1), sd, p + ".conv4")))      # :267-270
    return x, t_desc, t_feat


def inside_mask(boxes: torch.Tensor, src_hw, dst_hw) -> torch.Tensor:
    """a4. (N,4) clamped XYXY fp32 -> (N, h*w) float 0/1. utils.py:53-89, exact op order in fp32:
    scale by fp32(dst/src), centre=(a+b)*0.5, size=b-a, |c-p|/s <= 0.5 on integer pixel coords."""
    (sh, sw), (dh, dw) = src_hw, dst_hw
    b = boxes.to(torch.float32)
    rh = torch.tensor(dh / sh, dtype=torch.float32)
    rw = torch.tensor(dw / sw, dtype=torch.float32)
    x1, y1, x2, y2 = b[:, 0] * rw, b[:, 1] * rh, b[:, 2] * rw, b[:, 3] * rh
    xc, yc = (x1 + x2) * 0.5, (y1 + y2) * 0.5
    ws, hs = x2 - x1, y2 - y1
    ys = torch.arange(dh, dtype=torch.float32)
    xs = torch.arange(dw, dtype=torch.float32)
    in_y = (yc[:, None] - ys[None, :]).abs() / hs[:, None] <= 0.5          # (N,h)
    in_x = (xc[:, None] - xs[None, :]).abs() / ws[:, None] <= 0.5          # (N,w)
    return (in_y[:, :, None] & in_x[:, None, :]).flatten(1).float()


def _default_rasterizer():
    """detectron2's polygons_to_bitmask when it is installed, else the deterministic stand-in of lgd_b200.synth (the
    same choice the engine makes, lgd_b200.engine.polygon_rasterizer)."""
    try:
        from detectron2.structures.masks import polygons_to_bitmask  # type: ignore
        return polygons_to_bitmask
    except Exception:  # noqa: BLE001
        from lgd_b200.synth import polygons_to_bitmask
        return polygons_to_bitmask


def seg_inside_masks(hws, batched_inputs, src_hw, add_bg_box: bool, rasterizer):
    """get_segmask_inside_gt (dynamic_teacher/utils.py:92-132): per image the polygon masks rasterised at the image's
    own resolution (`rasterizer` = detectron2's polygons_to_bitmask), a background row of ones over the (un-padded)
    image when add_bg_box, zero padding to the padded batch size, then F.interpolate(mode='nearest') to every level.
    Returns F x B x (N_i, h*w) float masks."""
    per_img = []
    for item in batched_inputs:
        inst = item["instances"]
        _, H, W = item["image"].shape
        n = len(inst)
        box_size = max(n + (1 if add_bg_box else 0), 1)
        target = torch.zeros(box_size, H * W)
        label_idx = -1
        if n > 0:
            for label_idx, polys in enumerate(inst.gt_masks):
                m = torch.from_numpy(rasterizer(polys, H, W)).reshape(-1)
                target[label_idx, m] = 1.0
        if add_bg_box:
            target[label_idx + 1, :] = 1.0
        target = target.reshape(1, box_size, H, W).float()
        target = F.pad(target, (0, src_hw[1] - W, 0, src_hw[0] - H))
        per_img.append([F.interpolate(target, size=(h, w), mode="nearest").squeeze(0).flatten(1) for h, w in hws])
    return [list(x) for x in zip(*per_img)]


class _ConvTF32(torch.autograd.Function):
    """conv3x3 whose MMA operands are rounded to TF32 (rna) in forward, dgrad and wgrad -- an
    emulation of what the tcgen05 kind::tf32 kernels compute, used to predict parity margins."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return F.conv2d(round_tf32(x), round_tf32(w), b, stride=1, padding=1)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gr = round_tf32(g)
        gx = torch.nn.grad.conv2d_input(x.shape, round_tf32(w), gr, stride=1, padding=1)
        gw = torch.nn.grad.conv2d_weight(round_tf32(x), w.shape, gr, stride=1, padding=1)
        return gx, gw, g.sum((0, 2, 3))


def _conv(x, sd, name, tf32=False):
    w, b = sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype)
    if tf32:
        return _ConvTF32.apply(x, w, b)
    return F.conv2d(x, w, b, stride=1, padding=1)


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away, like cvt.rna.tf32.f32) emulation of TF32 operand rounding."""
    if x.dtype != torch.float32:
        x32 = x.to(torch.float32)
    else:
        x32 = x
    i = x32.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    r = torch.where(torch.isfinite(x32), r, x32)
    return r.to(x.dtype)


def _gn1(x):  # GroupNorm(1 group, no affine): per-sample over (C,H,W), layers.py:6-7
    return F.group_norm(x, 1, eps=EPS)


def _relu_site(x, ctl, site):
    """ReLU of one of the six pyramid-sized ReLU sites of the path (per level key k: "sp/k" student_proj GN-ReLU,
    "y0/k" local_inst_proj_2D (+ctx) ReLU, "y1/k" / "y2/k" refinement GN-ReLUs, "a1/k" / "a2/k" adapter ReLUs).
    ctl = None: plain F.relu (the reference). ctl = {"record": {}} stores the activation pattern x > 0 per site ("record_x": {} the pre-activations too);
    ctl = {"force": {site: bool tensor}} evaluates y = x * pattern with a GIVEN activation pattern instead of x > 0 --
    the gradient-parity tests run the oracle once with the pattern the CUDA engine took, which separates kernel
    accuracy from the discrete sign decisions of activations that are zero to within operand rounding."""
    if ctl is None:
        return F.relu(x)
    if "record" in ctl:
        ctl["record"][site] = (x > 0)
    if "record_x" in ctl:
        ctl["record_x"][site] = x.detach()
    force = ctl.get("force")
    if force is not None and site in force:
        return x * force[site].to(x.dtype)
    return F.relu(x)


def mha(query, kv, mask_img_q, mask_img_k, sd, heads, prefix="teacher.multi_head_attn"):
    """a6. nn.MultiheadAttention(256, heads), batch 1, block-diagonal mask (True = other image).
    query (Tq,256), kv (Tk,256). dynamic_teacher.py:255-275."""
    E = query.shape[1]
    hd = E // heads
    Wi, bi = sd[prefix + ".in_proj_weight"].to(query.dtype), sd[prefix + ".in_proj_bias"].to(query.dtype)
    q = F.linear(query, Wi[:E], bi[:E])
    k = F.linear(kv, Wi[E:2 * E], bi[E:2 * E])
    v = F.linear(kv, Wi[2 * E:], bi[2 * E:])
    q = q.view(-1, heads, hd).transpose(0, 1) * (hd ** -0.5)
    k = k.view(-1, heads, hd).transpose(0, 1)
    v = v.view(-1, heads, hd).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    ignore = mask_img_q[:, None] != mask_img_k[None, :]
    s = s.masked_fill(ignore[None], float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).reshape(-1, E)
    return F.linear(o, sd[prefix + ".out_proj.weight"].to(o.dtype), sd[prefix + ".out_proj.bias"].to(o.dtype))


# ----------------------------------------------------------------------------- full step
def teacher_forward(sd, instances, img_hw, features: Dict[str, torch.Tensor], *, add_context_box=True,
                    detach_appearance_embed=False, interact_pattern="stuGuided", heads=8,
                    dtype=torch.float32, tf32=False, keep=False, relu_ctl=None, box_format="x1y1x2y2",
                    seg=None, category_format="one_hot", exact_local_inst=False):
    """seg (LOAD_LABELMAP = True, the Mask R-CNN recipe): dict(batched_inputs=..., rasterizer=polygons_to_bitmask) --
    descriptors get the 49 mask dimensions and pooling / rendering use the rasterised polygon masks
    (dynamic_teacher.py:238-239) instead of the box masks.
    DynamicTeacher.forward (dynamic_teacher.py:285-301). Returns (features_tea dict, inst_labels,
    masks[F][B], stages dict). relu_ctl: see _relu_site (None = the reference's plain ReLUs)."""
    img_h, img_w = img_hw
    st = {}
    per_img = prepare_boxes(instances, img_h, img_w, add_context_box, box_format, add_mask=seg is not None,
                            category_format=category_format)
    counts = [p[0].shape[0] for p in per_img]
    desc = torch.cat([encode_descriptors(p[0], p[1], img_h, img_w, p[3] if seg is not None else None) for p in per_img], 0)
    seg_masks = None
    if seg is not None:
        hws_all = [tuple(features[k].shape[-2:]) for k in features]
        seg_masks = seg_inside_masks(hws_all, seg["batched_inputs"], (img_h, img_w), add_context_box, seg["rasterizer"])
    st["desc"] = desc
    sdd = sd
    label_embed, _, _ = label_encoder(desc.to(dtype), counts, sdd)
    st["label_embed"] = label_embed
    canoni = F.relu(_ln(_lin(label_embed, sdd, "teacher.canoni_proj_1D.0.0")))
    st["canoni"] = canoni
    img_of = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(counts)])
    keys = list(features.keys())
    B = len(counts)
    masks, attn_out, tea = [], [], {}
    st["stu_proj"], st["pooled"], st["attn"], st["rendered"], st["inst_map"] = [], [], [], [], []
    for lvl, key in enumerate(keys):
        x = features[key].to(dtype)
        if detach_appearance_embed:
            x = x.detach()
        _, _, h, w = x.shape
        proj = _relu_site(_gn1(_conv(x, sdd, "teacher.student_proj_2D.0.0", tf32)), relu_ctl, "sp/" + key)   # a3
        if seg_masks is not None:
            m_lvl = seg_masks[lvl]
        else:
            m_lvl = [inside_mask(p[0], (img_h, img_w), (h, w)) for p in per_img]            # a4
        masks.append(m_lvl)
        pooled = []
        for bi in range(B):                                                                # a5
            m = m_lvl[bi].to(dtype)
            cnt = torch.clamp(m.sum(-1), min=1.0)
            pooled.append((m @ proj[bi].flatten(1).T) / cnt[:, None])
        pooled = torch.cat(pooled, 0)
        if interact_pattern == "stuGuided":                                                # a6
            a = mha(pooled, canoni, img_of, img_of, sdd, heads)
        elif interact_pattern == "labelGuided":
            a = mha(canoni, pooled, img_of, img_of, sdd, heads)
        elif interact_pattern == "student_fill":
            a = pooled
        elif interact_pattern == "teacher_fill":
            a = canoni
        else:
            raise ValueError("interact pattern: {} not supported !".format(interact_pattern))
        # a7 rendering
        rendered, ctx_rows = [], []
        off = 0
        for bi in range(B):
            rows = a[off:off + counts[bi]]
            m = m_lvl[bi].to(dtype)
            off += counts[bi]
            if add_context_box:
                inst = _lin(rows[:-1], sdd, "teacher.local_inst_proj_1D")
                ctx_rows.append(rows[-1])
                rendered.append((inst.T @ m[:-1]).view(-1, h, w))
            else:
                inst = _lin(rows, sdd, "teacher.local_inst_proj_1D")
                rendered.append((inst.T @ m).view(-1, h, w))
        rendered = torch.stack(rendered, 0)
        # exact_local_inst: the engine evaluates this convolution in exact fp32 from per-box tap vectors (taprender.cu), so
        # the operand-rounding emulation leaves it un-rounded
        inst_map = _conv(rendered, sdd, "teacher.local_inst_proj_2D", tf32 and not exact_local_inst)
        if add_context_box:
            ctx = _lin(torch.stack(ctx_rows, 0), sdd, "teacher.global_ctx_proj_1D")
            y = _relu_site(inst_map + ctx[:, :, None, None], relu_ctl, "y0/" + key)
        else:
            y = _relu_site(inst_map, relu_ctl, "y0/" + key)
        # a8 refinement
        y = _relu_site(_gn1(_conv(y, sdd, "teacher.refinement_module.0", tf32)), relu_ctl, "y1/" + key)
        y = _relu_site(_gn1(_conv(y, sdd, "teacher.refinement_module.3", tf32)), relu_ctl, "y2/" + key)
        y = _gn1(_conv(y, sdd, "teacher.refinement_module.6", tf32))
        tea[key] = y
        if keep:
            st["stu_proj"].append(proj); st["pooled"].append(pooled); st["attn"].append(a)
            st["rendered"].append(rendered); st["inst_map"].append(inst_map)
    inst_labels = [p[2] for p in per_img]
    return tea, inst_labels, masks, st


def adapter_forward(sd, x, tf32=False, prefix="adapter.distill.adapter", relu_ctl=None, key=""):
    """a10. conv-ReLU-conv-ReLU-conv (sequential_convs.py:11-15)."""
    x = _relu_site(_conv(x, sd, prefix + ".0", tf32), relu_ctl, "a1/" + key)
    x = _relu_site(_conv(x, sd, prefix + ".2", tf32), relu_ctl, "a2/" + key)
    return _conv(x, sd, prefix + ".4", tf32)


def distill_loss(sd, stu: Dict[str, torch.Tensor], tea: Dict[str, torch.Tensor], lam=1.0,
                 distill_flag=1, dtype=torch.float32, tf32=False, relu_ctl=None):
    """a11. base_distillator.py:34-64: InstanceNorm both sides, MSE over all levels concatenated."""
    keys = sorted(stu.keys() & tea.keys())
    bs = tea[keys[0]].shape[0]
    s_list, t_list = [], []
    for k in keys:
        s = stu[k].to(dtype)
        if distill_flag == 0:
            s = s.detach()
        t = tea[k].detach().to(dtype)
        s = adapter_forward(sd, s, tf32, relu_ctl=relu_ctl, key=k)
        s_list.append(F.instance_norm(s, eps=EPS).reshape(bs, -1))
        t_list.append(F.instance_norm(t, eps=EPS).reshape(bs, -1))
    return lam * F.mse_loss(torch.cat(t_list, 1), torch.cat(s_list, 1))


def distill_step(sd, batched_inputs, images, features, *, add_context_box=True,
                 detach_appearance_embed=False, interact_pattern="stuGuided", heads=8, lam=1.0,
                 distill_flag=1, dtype=torch.float32, tf32=False, keep=False, relu_ctl=None, box_format="x1y1x2y2",
                 load_labelmap=False, rasterizer=None, category_format="one_hot", exact_local_inst=False):
    """teacher.forward -> distill_loss, as Distillator*.forward drives them (distillator.py:57-69)."""
    instances = [x["instances"] for x in batched_inputs]
    _, _, H, W = images.tensor.size()
    sd = {k: v.to(dtype) for k, v in sd.items()}
    tea, inst_labels, masks, st = teacher_forward(
        sd, instances, (H, W), features, add_context_box=add_context_box,
        detach_appearance_embed=detach_appearance_embed, interact_pattern=interact_pattern,
        heads=heads, dtype=dtype, tf32=tf32, keep=keep, relu_ctl=relu_ctl, box_format=box_format,
        seg=dict(batched_inputs=batched_inputs, rasterizer=rasterizer or _default_rasterizer()) if load_labelmap else None,
        category_format=category_format, exact_local_inst=exact_local_inst)
    loss = distill_loss(sd, features, tea, lam, distill_flag, dtype, tf32, relu_ctl=relu_ctl)
    return tea, inst_labels, masks, loss, st


# ----------------------------------------------------------------------------- f1: student head on teacher features
def retinanet_head(sd, features: Sequence[torch.Tensor], num_anchors: int = 9, num_classes: int = 80, relu_ctl=None,
                   dtype=torch.float32):
    """SURVEY.md 8(f) rank 1. RetinaNetCT.predict without the anchors (customized_detectors/retinanet.py:36-45):
    `pred_logits, pred_anchor_deltas = self.head(features)` followed by permute_to_N_HWA_K (retinanet.py:13-22).
    `self.head` is detectron2 0.3's RetinaNetHead -- a third-party dependency that is NOT in /root/reference and not
    installable here (README.md:67 pins detectron2==0.3): its published algorithm is restated from its definition:
    cls_subnet / bbox_subnet = 4 x [Conv2d(256,256,3,1,1), ReLU] shared by all levels, cls_score = Conv2d(256, A*K, 3,
    1, 1), bbox_pred = Conv2d(256, A*4, 3, 1, 1). PARITY UNPINNED for this function (no reference-owned vector exists);
    it is plain torch.nn.functional arithmetic. sd: the head's state_dict names. Returns (logits, deltas) as lists of
    (N, Hi*Wi*A, K) / (N, Hi*Wi*A, 4)."""
    def permute_to_N_HWA_K(t, K):
        N, _, H, W = t.shape
        return t.view(N, -1, K, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, K)

    logits, deltas = [], []
    for l, x in enumerate(features):
        x = x.to(dtype)
        c = b = x
        for i in (0, 2, 4, 6):
            c = _relu_site(F.conv2d(c, sd["cls_subnet.%d.weight" % i].to(dtype), sd["cls_subnet.%d.bias" % i].to(dtype),
                                    padding=1), relu_ctl, "cls%d/%d" % (i, l))
            b = _relu_site(F.conv2d(b, sd["bbox_subnet.%d.weight" % i].to(dtype), sd["bbox_subnet.%d.bias" % i].to(dtype),
                                    padding=1), relu_ctl, "box%d/%d" % (i, l))
        logits.append(permute_to_N_HWA_K(F.conv2d(c, sd["cls_score.weight"].to(dtype), sd["cls_score.bias"].to(dtype),
                                                  padding=1), num_classes))
        deltas.append(permute_to_N_HWA_K(F.conv2d(b, sd["bbox_pred.weight"].to(dtype), sd["bbox_pred.bias"].to(dtype),
                                                  padding=1), 4))
    return logits, deltas


TOWER_CONV = (0, 3, 6, 9)    # conv indices inside the FCOS-family towers: [Conv2d, GroupNorm(32), ReLU] x 4


def fcos_head(sd, features: Sequence[torch.Tensor], fpn_strides: Sequence[int], centerness_on_reg: bool = True,
              norm_reg_targets: bool = True, relu_ctl=None, dtype=torch.float32):
    """SURVEY.md 8(f) rank 1, FCOS family. FCOSHead.forward (thirdparty_heads/fcos.py:503-546; ATSS uses the same class,
    atss.py:97) and POTOHead.forward (poto.py:592-625: the same without the centerness branch -- selected by the absence
    of 'centerness.weight' in sd): per level
        cls_subnet / bbox_subnet = 4 x [Conv2d(256,256,3,1,1), GroupNorm(32,256) (affine), ReLU]      (fcos.py:455-476)
        logits     = cls_score(cls_subnet(x))                                                          (fcos.py:532)
        centerness = centerness(bbox_subnet(x) if centerness_on_reg else cls_subnet(x))                (fcos.py:533-536)
        bbox_pred  = scales[level](bbox_pred(bbox_subnet(x)))                                          (fcos.py:538)
        bbox_reg   = relu(bbox_pred) * fpn_strides[level] if norm_reg_targets else exp(bbox_pred)      (fcos.py:539-542)
    Pinned against the unmodified reference classes by tests/golden/fcos_head_*.npz (oracle/make_golden.py).
    sd: the head's state_dict names. relu_ctl: see _relu_site (sites "cls<i>/<level>", "box<i>/<level>", and "reg/<level>"
    for the ReLU of the box decoding).
    Returns (logits, bbox_reg, centerness) as lists of NCHW tensors (centerness = None for the POTO head)."""
    has_ctr = "centerness.weight" in sd
    logits, bbox_reg, centerness = [], [], []
    P = {k: v.to(dtype) for k, v in sd.items()}
    for l, x in enumerate(features):
        x = x.to(dtype)
        c = b = x
        for i in TOWER_CONV:
            c = F.conv2d(c, P["cls_subnet.%d.weight" % i], P["cls_subnet.%d.bias" % i], padding=1)
            c = F.group_norm(c, 32, P["cls_subnet.%d.weight" % (i + 1)], P["cls_subnet.%d.bias" % (i + 1)], 1e-5)
            c = _relu_site(c, relu_ctl, "cls%d/%d" % (i, l))
            b = F.conv2d(b, P["bbox_subnet.%d.weight" % i], P["bbox_subnet.%d.bias" % i], padding=1)
            b = F.group_norm(b, 32, P["bbox_subnet.%d.weight" % (i + 1)], P["bbox_subnet.%d.bias" % (i + 1)], 1e-5)
            b = _relu_site(b, relu_ctl, "box%d/%d" % (i, l))
        logits.append(F.conv2d(c, P["cls_score.weight"], P["cls_score.bias"], padding=1))
        if has_ctr:
            centerness.append(F.conv2d(b if centerness_on_reg else c, P["centerness.weight"], P["centerness.bias"], padding=1))
        pred = F.conv2d(b, P["bbox_pred.weight"], P["bbox_pred.bias"], padding=1) * P["scales.%d.scale" % l]
        bbox_reg.append(_relu_site(pred, relu_ctl, "reg/%d" % l) * fpn_strides[l] if norm_reg_targets else torch.exp(pred))
    return logits, bbox_reg, (centerness if has_ctr else None)
