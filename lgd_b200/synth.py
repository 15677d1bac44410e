"""Synthetic COCO-shaped inputs for the LGD distillation step, plus the tiny
stand-ins for the detectron2 structures the hot path touches.

The hot path reads only (reference citations):
  * ``batched_inputs[i]['instances']``: ``len()``, ``.gt_boxes.tensor`` (n,4) XYXY abs fp32,
    ``.gt_boxes.device``, ``.gt_classes`` (n,) int64   (label_encoder.py:40-51)
  * ``images.tensor.size()``  -> padded batch H, W        (label_encoder.py:166-167)
  * ``features`` dict p3..p7 of (B,256,H_l,W_l) fp32      (dynamic_teacher.py:209-235)

Workload definition follows SURVEY.md section 8(d): seeds, box statistics, level sizes.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
from typing import Dict, List, Sequence, Tuple

import torch

LEVEL_KEYS = ("p3", "p4", "p5", "p6", "p7")
NUM_CLASSES = 80
CHANNELS = 256


class Boxes:
    """Stand-in for detectron2.structures.Boxes (only .tensor / .device are read)."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor

    @property
    def device(self):
        return self.tensor.device


def polygons_to_bitmask(polygons, height: int, width: int) -> np.ndarray:
    """Stand-in for detectron2.structures.masks.polygons_to_bitmask (pycocotools is not installed here): union of the
    instance's polygons, even-odd rule evaluated at pixel centres. `polygons`: list of flat [x0, y0, x1, y1, ...]
    arrays. Any deterministic rasteriser serves the parity tests: the reference (through oracle/refshim.py) and the
    engine are handed the same function, exactly as both would call pycocotools' one in production."""
    ys, xs = np.mgrid[0:height, 0:width]
    px, py = xs + 0.5, ys + 0.5
    out = np.zeros((height, width), dtype=bool)
    for poly in polygons:
        pts = np.asarray(poly, dtype=np.float64).reshape(-1, 2)
        inside = np.zeros((height, width), dtype=bool)
        n = len(pts)
        for i in range(n):
            x0, y0 = pts[i]
            x1, y1 = pts[(i + 1) % n]
            if y0 == y1:
                continue
            cond = (y0 > py) != (y1 > py)
            xint = (x1 - x0) * (py - y0) / (y1 - y0) + x0
            inside ^= cond & (px < xint)
        out |= inside
    return out


class PolygonMasks:
    """Stand-in for detectron2.structures.PolygonMasks: iteration yields one polygon list per instance
    (dynamic_teacher/utils.py:112-114), crop_and_resize(boxes, M) the (N, M, M) box-relative bitmasks of the mask
    descriptors (label_encoder.py:60-63)."""

    def __init__(self, polygons):
        self.polygons = polygons

    def __len__(self):
        return len(self.polygons)

    def __iter__(self):
        return iter(self.polygons)

    def crop_and_resize(self, boxes: torch.Tensor, mask_size: int) -> torch.Tensor:
        out = torch.zeros(len(self.polygons), mask_size, mask_size, dtype=torch.bool)
        for i, (polys, box) in enumerate(zip(self.polygons, boxes.tolist())):
            x1, y1, x2, y2 = box
            w, h = max(x2 - x1, 1e-6), max(y2 - y1, 1e-6)
            shifted = [(np.asarray(p, dtype=np.float64).reshape(-1, 2) - [x1, y1]) * [mask_size / w, mask_size / h] for p in polys]
            out[i] = torch.from_numpy(polygons_to_bitmask([q.reshape(-1) for q in shifted], mask_size, mask_size))
        return out


class Instances:
    """Stand-in for detectron2.structures.Instances."""

    def __init__(self, gt_boxes: torch.Tensor, gt_classes: torch.Tensor, gt_masks=None):
        self.gt_boxes = Boxes(gt_boxes)
        self.gt_classes = gt_classes
        if gt_masks is not None:
            self.gt_masks = gt_masks

    def __len__(self):
        return int(self.gt_boxes.tensor.shape[0])


class ImageList:
    """Stand-in for detectron2.structures.ImageList (only .tensor.size() is read)."""

    def __init__(self, tensor: torch.Tensor, image_sizes=None):
        self.tensor = tensor
        self.image_sizes = image_sizes


def make_cfg(
    add_context_box: bool = True,
    detach_appearance_embed: bool = False,
    interact_pattern: str = "stuGuided",
    lam: float = 1.0,
    device: str = "cpu",
    heads: int = 8,
    box_format: str = "x1y1x2y2",
    load_labelmap: bool = False,
    category_format: str = "one_hot",
):
    """Attribute bag with the cfg keys the hot path reads (utils/build.py:557-653)."""
    ns = SimpleNamespace
    cfg = ns(
        NUM_CLASSES=NUM_CLASSES,
        MODEL=ns(
            DEVICE=device,
            FPN=ns(OUT_CHANNELS=CHANNELS),
            RECIPROCAL_FPN_STRIDES=[1 / 8, 1 / 16, 1 / 32, 1 / 64, 1 / 128],
            DISTILLATOR=ns(
                LAMBDA=lam,
                ADAPTER=ns(META_ARCH="SequentialConvs"),
                LABEL_ENCODER=ns(
                    BOX_FORMAT=box_format, CATEGORY_FORMAT=category_format, LOAD_LABELMAP=load_labelmap
                ),
                TEACHER=ns(
                    META_ARCH="DynamicTeacher",
                    INTERACT_PATTERN=interact_pattern,
                    ADD_CONTEXT_BOX=add_context_box,
                    DETACH_APPEARANCE_EMBED=detach_appearance_embed,
                    NR_TRANSFORMER_HEADS=heads,
                ),
                STUDENT=ns(META_ARCH=None),
            ),
        ),
    )
    return cfg


def pyramid_hw(img_h: int, img_w: int) -> List[Tuple[int, int]]:
    """p3..p7 sizes for a padded image: stride 8/16/32, then two stride-2 3x3 convs (ceil)."""
    h, w = img_h // 8, img_w // 8
    out = [(h, w)]
    for _ in range(2):
        h, w = h // 2, w // 2
        out.append((h, w))
    for _ in range(2):
        h, w = (h + 1) // 2, (w + 1) // 2
        out.append((h, w))
    return out


def pad32(n: int) -> int:
    return (n + 31) // 32 * 32


def synth_boxes(gen: torch.Generator, img_h: int, img_w: int, mean_boxes: float = 7.3,
                max_boxes: int = 64, n: int | None = None):
    """One image worth of GT: n ~ clamp(round(Exp(mean)),1,max); log-uniform sizes 8px..full."""
    if n is None:
        u = torch.rand((), generator=gen).clamp_min(1e-12)
        n = int(min(max(round(float(-mean_boxes * math.log(float(u)))), 1), max_boxes))
    cx = torch.rand(n, generator=gen) * img_w
    cy = torch.rand(n, generator=gen) * img_h
    bw = 8.0 * (img_w / 8.0) ** torch.rand(n, generator=gen)
    bh = 8.0 * (img_h / 8.0) ** torch.rand(n, generator=gen)
    x1 = (cx - bw / 2).clamp(0, img_w)
    x2 = (cx + bw / 2).clamp(0, img_w)
    y1 = (cy - bh / 2).clamp(0, img_h)
    y2 = (cy + bh / 2).clamp(0, img_h)
    boxes = torch.stack([x1, y1, x2, y2], dim=1).float()
    classes = torch.randint(0, NUM_CLASSES, (n,), generator=gen, dtype=torch.int64)
    return boxes, classes


def adversarial_boxes(img_h: int, img_w: int):
    """Edge cases of SURVEY 8(d): stride-aligned integer coords, zero-width/zero-height boxes,
    boxes touching / exceeding the border, one-pixel boxes."""
    b = [
        [0.0, 0.0, float(img_w), float(img_h)],          # whole image (gets clamped to w-1,h-1)
        [8.0, 16.0, 64.0, 128.0],                        # stride aligned
        [32.0, 32.0, 32.0, 96.0],                        # zero width
        [40.0, 48.0, 120.0, 48.0],                       # zero height
        [float(img_w) - 1.0, float(img_h) - 1.0, float(img_w) + 5.0, float(img_h) + 7.0],  # corner
        [16.0, 16.0, 17.0, 17.0],                        # one pixel
        [0.5, 0.5, 7.5, 7.5],                            # sub-stride
        [-5.0, -3.0, 30.0, 20.0],                        # negative coords (clamped)
    ]
    boxes = torch.tensor(b, dtype=torch.float32)
    classes = torch.tensor([0, 79, 1, 2, 3, 40, 41, 5], dtype=torch.int64)
    return boxes, classes


def synth_batch(B: int, img_h: int = 800, img_w: int = 1333, seed: int = 1234, device="cpu",
                n_boxes: Sequence[int | None] | None = None, adversarial: bool = False,
                level_keys: Sequence[str] = LEVEL_KEYS, feature_device=None, requires_grad=False,
                with_masks: bool = False, unpadded: Sequence[Tuple[int, int]] | None = None):
    """Returns (batched_inputs, images, features) shaped like what Distillator*.forward hands to
    the teacher (distillator.py:96-104). Feature maps are i.i.d. N(0,1) fp32 NCHW."""
    gen = torch.Generator().manual_seed(seed)
    H, W = pad32(img_h), pad32(img_w)
    hws = pyramid_hw(H, W)
    batched_inputs = []
    for i in range(B):
        if adversarial and i == 0:
            boxes, classes = adversarial_boxes(img_h, img_w)
        else:
            n = None if n_boxes is None else n_boxes[i]
            if n == 0:
                boxes = torch.zeros(0, 4)
                classes = torch.zeros(0, dtype=torch.int64)
            else:
                boxes, classes = synth_boxes(gen, img_h, img_w, n=n)
        item = {"height": img_h, "width": img_w}
        if with_masks:
            # polygon masks (Mask R-CNN recipe, LOAD_LABELMAP): an inscribed octagon plus, for every second box, a
            # triangle -- two polygons of one instance, overlapping neighbours; "image" carries the un-padded size
            ih, iw = unpadded[i] if unpadded is not None else (img_h, img_w)
            polys = []
            for j, (x1, y1, x2, y2) in enumerate(boxes.tolist()):
                cx, cy, rx, ry = (x1 + x2) / 2, (y1 + y2) / 2, max((x2 - x1) / 2, 0.5), max((y2 - y1) / 2, 0.5)
                octo = [c for a in range(8) for c in (cx + rx * math.cos(a * math.pi / 4 + 0.2), cy + ry * math.sin(a * math.pi / 4 + 0.2))]
                p = [np.asarray(octo)]
                if j % 2 == 1:
                    p.append(np.asarray([x1, y1, x2, y1 + 0.3 * (y2 - y1), x1 + 0.2 * (x2 - x1), y2]))
                polys.append(p)
            item["instances"] = Instances(boxes, classes, PolygonMasks(polys))
            item["image"] = torch.empty(3, ih, iw, device="meta")
        else:
            item["instances"] = Instances(boxes, classes)
        batched_inputs.append(item)
    images = ImageList(torch.empty(B, 3, H, W, device="meta"), [(img_h, img_w)] * B)
    fdev = feature_device if feature_device is not None else device
    features: Dict[str, torch.Tensor] = {}
    for key, (h, w) in zip(level_keys, hws):
        t = torch.randn(B, CHANNELS, h, w, generator=gen, dtype=torch.float32)
        t = t.to(fdev)
        if requires_grad:
            t.requires_grad_(True)
        features[key] = t
    return batched_inputs, images, features


def synth_cotangents(features: Dict[str, torch.Tensor], seed: int = 4321, sigma: float = 1e-3):
    """Fixed random cotangents standing in for the student-head gradient on the teacher pyramid
    (SURVEY 8(d) 'distillation step', fwd+bwd definition)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in features.items():
        out[k] = (torch.randn(v.shape, generator=gen, dtype=torch.float32) * sigma).to(v.device)
    return out


# --------------------------------------------------------------------------- deterministic weights
# Names/shapes = the reference's checkpoint contract (SURVEY.md 8(b) "state_dict names").
def desc_dim_of(cfg_kw) -> int:
    """Descriptor length of a configuration (label_encoder.py:24-32,136-145): 4 box coordinates + 80 one-hot classes, or
    + 1 normalised class index (CATEGORY_FORMAT norm_classes); + 49 mask dimensions with LOAD_LABELMAP."""
    d = 4 + (1 if cfg_kw.get("category_format", "one_hot") == "norm_classes" else NUM_CLASSES)
    return d + (49 if cfg_kw.get("load_labelmap") else 0)


def hot_path_param_shapes(desc_dim: int = 84):
    """desc_dim = 84 (boxes + one-hot classes), 133 (+ 49 mask dimensions, LOAD_LABELMAP) or 5 / 54 (norm_classes)"""
    shapes = {}

    def stn(p, k):
        shapes[p + ".conv1.weight"] = (64, k, 1); shapes[p + ".conv1.bias"] = (64,)
        shapes[p + ".conv2.weight"] = (128, 64, 1); shapes[p + ".conv2.bias"] = (128,)
        shapes[p + ".conv3.weight"] = (1024, 128, 1); shapes[p + ".conv3.bias"] = (1024,)
        shapes[p + ".fc1.weight"] = (512, 1024); shapes[p + ".fc1.bias"] = (512,)
        shapes[p + ".fc2.weight"] = (256, 512); shapes[p + ".fc2.bias"] = (256,)
        shapes[p + ".fc3.weight"] = (k * k, 256); shapes[p + ".fc3.bias"] = (k * k,)

    le = "teacher.label_encoder_"
    stn(le + ".stn_desc", desc_dim)
    stn(le + ".stn_feat", 64)
    for name, (o, i) in {"conv1": (64, desc_dim), "conv2": (128, 64), "conv3": (1024, 128), "conv4": (256, 1088)}.items():
        shapes[f"{le}.{name}.weight"] = (o, i, 1); shapes[f"{le}.{name}.bias"] = (o,)
    for name in ("teacher.canoni_proj_1D.0.0", "teacher.global_ctx_proj_1D", "teacher.local_inst_proj_1D"):
        shapes[name + ".weight"] = (256, 256); shapes[name + ".bias"] = (256,)
    for name in ("teacher.student_proj_2D.0.0", "teacher.local_inst_proj_2D", "teacher.refinement_module.0",
                 "teacher.refinement_module.3", "teacher.refinement_module.6",
                 "adapter.distill.adapter.0", "adapter.distill.adapter.2", "adapter.distill.adapter.4"):
        shapes[name + ".weight"] = (256, 256, 3, 3); shapes[name + ".bias"] = (256,)
    shapes["teacher.multi_head_attn.in_proj_weight"] = (768, 256)
    shapes["teacher.multi_head_attn.in_proj_bias"] = (768,)
    shapes["teacher.multi_head_attn.out_proj.weight"] = (256, 256)
    shapes["teacher.multi_head_attn.out_proj.bias"] = (256,)
    return shapes


def synth_state_dict(seed: int = 0, bias_scale: float = 1.0, desc_dim: int = 84):
    """Deterministic random weights with PyTorch-default-init statistics (U(+-1/sqrt(fan_in))),
    generated name by name from a seeded CPU generator so every implementation (reference,
    oracle, CUDA engine) can be loaded with bit-identical parameters."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in sorted(hot_path_param_shapes(desc_dim).items()):
        if name.endswith("weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
        else:
            fan_in = 256
        bound = 1.0 / math.sqrt(fan_in)
        t = (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound
        if name.endswith("bias"):
            t = t * bias_scale
        sd[name] = t
    return sd


# --------------------------------------------------------------------------- FCOS-family detection head (8(f) rank 1)
def fcos_head_param_shapes(num_levels: int = 5, num_classes: int = 80, centerness: bool = True):
    """state_dict names / shapes of FCOSHead (thirdparty_heads/fcos.py:438-501; POTOHead, poto.py:528-590, is the same
    without the centerness convolution): towers of [Conv2d(256,256,3), GroupNorm(32,256), ReLU] x 4."""
    shapes = {}
    for tower in ("cls_subnet", "bbox_subnet"):
        for i in (0, 3, 6, 9):
            shapes["%s.%d.weight" % (tower, i)] = (256, 256, 3, 3)
            shapes["%s.%d.bias" % (tower, i)] = (256,)
            shapes["%s.%d.weight" % (tower, i + 1)] = (256,)     # GroupNorm affine
            shapes["%s.%d.bias" % (tower, i + 1)] = (256,)
    shapes["cls_score.weight"], shapes["cls_score.bias"] = (num_classes, 256, 3, 3), (num_classes,)
    shapes["bbox_pred.weight"], shapes["bbox_pred.bias"] = (4, 256, 3, 3), (4,)
    if centerness:
        shapes["centerness.weight"], shapes["centerness.bias"] = (1, 256, 3, 3), (1,)
    for l in range(num_levels):
        shapes["scales.%d.scale" % l] = (1,)
    return shapes


def synth_fcos_head_state_dict(seed: int = 0, num_levels: int = 5, num_classes: int = 80, centerness: bool = True):
    """Deterministic head weights at a trained-network scale (the reference's N(0, 0.01) init would leave the towers
    numerically trivial): conv weights ~ N(0, 1/sqrt(fan_in)) so activations stay O(1) through the towers, GroupNorm gains
    around 1, shifts and biases O(0.1), per-level scales around 1."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in fcos_head_param_shapes(num_levels, num_classes, centerness).items():
        if len(shape) == 4:
            sd[name] = torch.randn(shape, generator=gen) / (shape[1] * 9) ** 0.5
        elif name.endswith(".scale"):
            sd[name] = 1.0 + 0.2 * torch.randn(shape, generator=gen)
        elif name.split(".")[1] in ("1", "4", "7", "10") and name.endswith("weight"):
            sd[name] = 1.0 + 0.2 * torch.randn(shape, generator=gen)
        else:
            sd[name] = 0.1 * torch.randn(shape, generator=gen)
    return sd
