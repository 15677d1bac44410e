"""the student's RetinaNet head on the teacher pyramid (lgd_b200/heads.py) against the CPU
restatement of detectron2's RetinaNetHead + permute_to_N_HWA_K (oracle.lgd_oracle.retinanet_head): forward within 1e-3;
gradients of the output convolutions (above every ReLU) within 1e-3, everything else against the oracle evaluated with
the engine's activation pattern (oracle/parity.py explains why) within 2e-3; and the zero-copy hand-over of the input
gradient to the teacher backward."""
import pytest
import torch
import torch.nn as nn

from lgd_b200 import synth
from lgd_b200.heads import PARAM_NAMES, RetinaNetHeadB200
from oracle import lgd_oracle as O

pytestmark = pytest.mark.gpu

A, K = 9, 80


def _make_head(seed=3):
    torch.manual_seed(seed)
    def tower():
        return nn.Sequential(*[m for _ in range(4) for m in (nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU())])
    head = nn.Module()
    head.cls_subnet, head.bbox_subnet = tower(), tower()
    head.cls_score = nn.Conv2d(256, A * K, 3, 1, 1)
    head.bbox_pred = nn.Conv2d(256, A * 4, 3, 1, 1)
    return head


def _rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,hw", [(2, (96, 128)), (1, (200, 264))])
def test_retinanet_head_matches_oracle(B, hw):
    from lgd_b200 import engine
    head = _make_head().cuda()
    b200 = RetinaNetHeadB200.from_module(head)
    assert RetinaNetHeadB200.supports(head)
    sd = {k: v.detach().cpu() for k, v in head.state_dict().items()}
    hws = synth.pyramid_hw(synth.pad32(hw[0]), synth.pad32(hw[1]))
    g = engine.Geometry.get(B, hws, torch.device("cuda", 0))
    gen = torch.Generator().manual_seed(11)
    pyr = torch.randn(g.elems, generator=gen).cuda().requires_grad_(True)     # an NHWC pyramid buffer like the teacher's
    feats = g.level_views(pyr)
    logits, deltas = b200(feats)
    # the loss side concatenates the levels per image (detectron2 losses): gradients arrive as strided slices
    cl, cd = torch.cat(logits, 1), torch.cat(deltas, 1)
    G1 = (torch.randn(cl.shape, generator=gen) * 1e-3).cuda()
    G2 = (torch.randn(cd.shape, generator=gen) * 1e-3).cuda()
    ((cl * G1).sum() + (cd * G2).sum()).backward()
    torch.cuda.synchronize()
    S = logits[0].grad_fn.S
    # activation patterns of the eight tower ReLUs as the engine took them (fp16 copies: nonzero = pass)
    force = {}
    for tower, tag in (("cls_subnet", "cls"), ("bbox_subnet", "box")):
        for k, i in enumerate((0, 2, 4, 6)):
            for l, v in enumerate(g.level_views(S.acts[tower][k + 1])):
                force["%s%d/%d" % (tag, i, l)] = (v != 0).cpu()
    fo = [f.detach().cpu().contiguous().requires_grad_(True) for f in feats]
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo, do = O.retinanet_head(sdo, fo, A, K, relu_ctl={"force": force})
    with torch.no_grad():
        lp, dp = O.retinanet_head(sd, [f.detach() for f in fo], A, K)
    for l in range(len(hws)):
        assert logits[l].shape == lp[l].shape and deltas[l].shape == dp[l].shape
        assert _rel(logits[l], lp[l]) < 1e-3, (l, _rel(logits[l], lp[l]))
        assert _rel(deltas[l], dp[l]) < 1e-3, (l, _rel(deltas[l], dp[l]))
    tot = (torch.cat(lo, 1) * G1.cpu()).sum() + (torch.cat(do, 1) * G2.cpu()).sum()
    names = sorted(sdo)
    grads = torch.autograd.grad(tot, fo + [sdo[n] for n in names])
    own = dict(b200.named_parameters())
    worst = {}
    for n, gr in zip(names, grads[len(fo):]):
        worst[n] = _rel(own[n].grad, gr)
    gx = pyr.grad
    for l, (v, gr) in enumerate(zip(g.level_views(gx), grads[:len(fo)])):
        worst["feat%d" % l] = _rel(v, gr)
    print("head gradient errors vs the pattern-evaluated oracle:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    for n in ("cls_score.weight", "cls_score.bias", "bbox_pred.weight", "bbox_pred.bias"):
        assert worst[n] < 1e-3, (n, worst[n])
    assert max(worst.values()) < 2e-3, sorted(worst.items(), key=lambda kv: -kv[1])[:5]


def test_head_backward_feeds_teacher_backward_in_place():
    """teacher pyramid -> head -> loss: the head's input gradient arrives at _TeacherFn.backward as level views of one
    NHWC buffer and is read in place (no NCHW round trip); the result equals the same step with the head run by plain
    PyTorch convolutions on the same device (fp32, TF32 off) within the step's gradient bars."""
    from tests.gpu_util import make_model
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(2, 120, 150, seed=31)
    head = _make_head(5).cuda()
    b200 = RetinaNetHeadB200.from_module(head)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    for mode in ("b200", "torch"):
        m = make_model(dict(add_context_box=True), sd, 1)
        head.zero_grad()
        f = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss = m.forward(bi, im, f)
        keys = list(tea.keys())
        if mode == "b200":
            logits, deltas = b200([tea[k] for k in keys])
        else:
            logits, deltas = [], []
            for k in keys:
                x = tea[k]
                n = x.shape[0]
                logits.append(head.cls_score(head.cls_subnet(x)).view(n, A, K, *x.shape[-2:]).permute(0, 3, 4, 1, 2).reshape(n, -1, K))
                deltas.append(head.bbox_pred(head.bbox_subnet(x)).view(n, A, 4, *x.shape[-2:]).permute(0, 3, 4, 1, 2).reshape(n, -1, 4))
        total = loss + (torch.cat(logits, 1) ** 2).mean() + (torch.cat(deltas, 1) ** 2).mean()
        total.backward()
        torch.cuda.synchronize()
        res[mode] = (float(total), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None},
                     {n: p.grad.clone() for n, p in head.named_parameters()}, {k: v.grad.clone() for k, v in f.items()})
    a, b = res["b200"], res["torch"]
    assert abs(a[0] - b[0]) <= 1e-3 * abs(b[0])
    for n in b[2]:
        assert _rel(a[2][n], b[2][n]) < 5e-2, (n, _rel(a[2][n], b[2][n]))
    for n in ("cls_score.weight", "bbox_pred.weight"):
        assert _rel(a[2][n], b[2][n]) < 2e-3, (n, _rel(a[2][n], b[2][n]))
    for n in b[1]:
        if n.endswith("adapter.4.bias"):
            continue
        assert _rel(a[1][n], b[1][n]) < 8e-2, (n, _rel(a[1][n], b[1][n]))
    for k in b[3]:
        assert _rel(a[3][k], b[3][k]) < 8e-2, k
