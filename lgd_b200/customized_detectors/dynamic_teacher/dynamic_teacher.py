"""DynamicTeacher: drop-in for models/customized_detectors/dynamic_teacher/dynamic_teacher.py:16-301.

Same registry name, constructor (`cls(cfg)`), call signature and return triple, same parameter names and
shapes (checkpoint compatible). The modules below only hold parameters; forward and backward run through
engine.teacher_forward / teacher_backward, i.e. hand-written sm_100a kernels behind the C ABI."""
import torch
import torch.nn as nn

from ... import engine
from ..build import CUSTOMIZED_DETECTORS_REGISTRY
from .label_encoder import LabelEncoder


class _TeacherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, batched_inputs, img_hw, names, n_feat, *tensors):
        feats, params = tensors[:n_feat], tensors[n_feat:]
        P = {"teacher." + n: p for n, p in zip(names, params)}
        mod._packed.new_step()
        with torch.cuda.device(feats[0].device):   # launches go to the current device's current stream
            # one native call per chain for the recipes the reference's configs train (one-hot classes, box masks);
            # LOAD_LABELMAP and CATEGORY_FORMAT norm_classes run the same kernels from the per-kernel orchestration
            if engine.chain_applicable(mod.interact_pattern) and not mod.use_seg_map \
                    and mod.category_format == 'one_hot':
                tea, S = engine.chain_teacher_forward(
                    P, feats, batched_inputs, img_hw, add_context_box=mod.add_context_box,
                    heads=mod.nr_transformer_heads, want_masks=mod.return_masks, box_format=mod.box_format)
            else:
                tea, S = engine.teacher_forward(
                    P, feats, batched_inputs, img_hw, add_context_box=mod.add_context_box,
                    interact_pattern=mod.interact_pattern, heads=mod.nr_transformer_heads, packed=mod._packed,
                    want_masks=mod.return_masks, box_format=mod.box_format, use_seg_map=mod.use_seg_map,
                    category_format=mod.category_format)
        # what distill() of the same step reuses (student operand pair, teacher pyramid buffer); it drops the cache
        # once it has consumed it, and the next forward overwrites it
        mod._step_cache = {"key": tuple((f.data_ptr(), f._version) for f in feats), "feats": feats, "stu": S.stu,
                           "stu_h": S.stu_h, "g": S.g, "tea": tea, "tea_stats": S.tea_in_stats,
                           "stu_ready": S.stu_ready}
        # the tape (about ten pyramid-sized buffers) lives in the autograd node only and is freed with it; the module
        # keeps it only on request (parity tests / diagnostics read intermediate stages from it)
        mod._last = S if mod.keep_tape else None
        mod._fwd_info = (S.tb, S.g, S.masks)
        ctx.mod, ctx.S, ctx.P, ctx.names, ctx.n_feat = mod, S, P, names, n_feat
        ctx.need_feat = (not mod.detach_appearance_embed) and any(f.requires_grad for f in feats)
        ctx.feat_needs = [f.requires_grad for f in feats]
        return tuple(S.g.level_views(tea))

    @staticmethod
    def backward(ctx, *gouts):
        S, g = ctx.S, ctx.S.g
        with torch.cuda.device(g.device):
            if getattr(S, "chain", False):
                grads, outs = engine.chain_teacher_backward(S, gouts, ctx.need_feat)
                gfeats = [None] * ctx.n_feat
                if outs is not None:
                    gfeats = [o if need else None for o, need in zip(outs, ctx.feat_needs)]
            else:
                gs = [go if go is not None else torch.zeros(g.B, 256, h, w, device=g.device)
                      for go, (h, w) in zip(gouts, g.hws)]
                g_tea = engine.to_pyramid(g, gs, False)
                grads, g_stu = engine.teacher_backward(ctx.P, S, g_tea, ctx.mod._packed, ctx.need_feat)
                gfeats = [None] * ctx.n_feat
                if g_stu is not None:
                    outs = engine.from_pyramid_nchw(g, g_stu)
                    gfeats = [o if need else None for o, need in zip(outs, ctx.feat_needs)]
        gparams = [grads.get("teacher." + n) for n in ctx.names]
        return (None, None, None, None, None, *gfeats, *gparams)


@CUSTOMIZED_DETECTORS_REGISTRY.register()
class DynamicTeacher(nn.Module):
    """dynamic teacher: (1) label encoder (2) inter-object relation adapter (3) intra-object knowledge mapper and the
    parameter-free appearance encoder (mask pooling)."""

    def __init__(self, cfg):
        super().__init__()
        self.nr_fpn_channels = cfg.MODEL.FPN.OUT_CHANNELS
        self.num_classes = cfg.NUM_CLASSES
        assert self.nr_fpn_channels == 256
        assert self.num_classes == 80
        self.interact_pattern = cfg.MODEL.DISTILLATOR.TEACHER.INTERACT_PATTERN
        self.strides = cfg.MODEL.RECIPROCAL_FPN_STRIDES
        self.box_format = cfg.MODEL.DISTILLATOR.LABEL_ENCODER.BOX_FORMAT
        self.category_format = cfg.MODEL.DISTILLATOR.LABEL_ENCODER.CATEGORY_FORMAT
        self.use_seg_map = cfg.MODEL.DISTILLATOR.LABEL_ENCODER.LOAD_LABELMAP
        self.add_context_box = cfg.MODEL.DISTILLATOR.TEACHER.ADD_CONTEXT_BOX
        self.detach_appearance_embed = cfg.MODEL.DISTILLATOR.TEACHER.DETACH_APPEARANCE_EMBED
        # construction order = the reference's, so default initialisation under a seed matches it
        self.label_encoder_ = LabelEncoder(category_format=self.category_format, box_format=self.box_format,
                                           nr_fg_classes=self.num_classes, add_context_box=self.add_context_box,
                                           parse_mask=self.use_seg_map)
        self.render_divide_occurence = False
        self.affine_flag = False
        c = self.nr_fpn_channels
        self.canoni_proj_1D = nn.Sequential(nn.Sequential(nn.Linear(c, c)))           # + LayerNorm + ReLU (no params)
        self.student_proj_2D = nn.Sequential(nn.Sequential(nn.Conv2d(c, c, 3, 1, 1)))  # + GroupNorm(1) + ReLU
        self.local_inst_proj_2D = nn.Conv2d(c, c, 3, 1, 1)
        self.global_ctx_proj_1D = nn.Linear(c, c)
        self.local_inst_proj_1D = nn.Linear(c, c)
        self.refinement_module = nn.Sequential(
            nn.Conv2d(c, c, 3, 1, 1), nn.Identity(), nn.Identity(),   # conv, GN(1), ReLU
            nn.Conv2d(c, c, 3, 1, 1), nn.Identity(), nn.Identity(),
            nn.Conv2d(c, c, 3, 1, 1), nn.Identity())
        self.nr_transformer_heads = cfg.MODEL.DISTILLATOR.TEACHER.NR_TRANSFORMER_HEADS
        self.multi_head_attn = nn.MultiheadAttention(c, self.nr_transformer_heads)
        self.return_masks = True   # the reference returns the float masks; set False to skip materialising them
        self._packed = engine.PackedWeights()
        self._step_cache = None
        self.keep_tape = False   # True: keep the last forward's tape in self._last (diagnostics; several GB at B=16)
        self._last = None
        self._fwd_info = None

    def forward(self, info_list):
        """info_list = (batched_inputs, images, r_features, features); returns
        (interact_tea_feats: dict, inst_labels: B x (Ni,), batchified_inside_masks: F x B x (Ni, HiWi))."""
        batched_inputs, images, features = info_list[0], info_list[1], info_list[-1]
        if self.interact_pattern not in ('stuGuided', 'labelGuided', 'student_fill', 'teacher_fill'):
            raise ValueError('interact pattern: {} not supported !'.format(self.interact_pattern))
        _, _, h, w = images.tensor.size()
        keys = list(features.keys())
        feats = [features[k] for k in keys]
        if feats[0].device.type != 'cuda':
            raise RuntimeError('lgd_b200.DynamicTeacher runs on CUDA (sm_100a) only; there is no CPU fallback')
        named = list(self.named_parameters())
        names = tuple(n for n, _ in named)
        outs = _TeacherFn.apply(self, batched_inputs, (int(h), int(w)), names, len(feats), *feats,
                                *[p for _, p in named])
        tb, g, flat_masks = self._fwd_info
        self._fwd_info = None
        tea = {k: o for k, o in zip(keys, outs)}
        masks = []
        if flat_masks is not None:
            off = 0
            for (hh, ww) in g.hws:
                lvl = flat_masks[off:off + tb.T * hh * ww].view(tb.T, hh * ww)
                masks.append(list(lvl.split(tb.counts, dim=0)))
                off += tb.T * hh * ww
        return tea, tb.inst_labels, masks
