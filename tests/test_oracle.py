"""CPU: the oracle restatement (oracle/lgd_oracle.py) against the golden vectors produced by the
unmodified reference (oracle/make_golden.py). Masks bit-exact; floats to fp32 round-off."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lgd_b200 import synth
from oracle import lgd_oracle as O
from oracle.make_golden import CASES, HEAD_CASES
from tests.golden_util import load_case, load_head_case, rel_l2, unpack_mask

FP32_TOL = 2e-5   # same algorithm, same fp32 ops, different summation order inside torch kernels


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case(name)
    feats = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tea, inst_labels, masks, loss, st = O.distill_step(sdo, bi, im, feats, distill_flag=flag, keep=True, **cfg_kw)
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_l2(st["label_embed"].detach(), g["label_embed"]) < FP32_TOL
    assert rel_l2(st["canoni"].detach(), g["canoni"]) < FP32_TOL
    for l, k in enumerate(feats):
        assert torch.equal(torch.cat(masks[l], 0), unpack_mask(g, k)), "mask membership must be bit exact"
        assert rel_l2(tea[k].detach(), g[f"tea_{k}"]) < FP32_TOL
        if cfg_kw.get("interact_pattern", "stuGuided") == "stuGuided":
            assert rel_l2(st["pooled"][l].detach(), g[f"mha_q_{k}"]) < FP32_TOL
        assert rel_l2(st["attn"][l].detach(), g[f"mha_out_{k}"]) < FP32_TOL
    for i, il in enumerate(inst_labels):
        assert np.array_equal(il.numpy().astype(np.float32), g[f"inst_labels_{i}"])
    # backward: same total as make_golden (loss + <tea, cotangent>)
    cot = synth.synth_cotangents(tea)
    total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
    names = sorted(sdo)
    grads = torch.autograd.grad(total, list(feats.values()) + [sdo[n] for n in names], allow_unused=True)
    for l, k in enumerate(feats):
        ref = g[f"gfeat_{k}"]
        if ref.size == 0:
            assert grads[l] is None
        else:
            assert rel_l2(grads[l], ref) < 5e-4, k
    for n, gr in zip(names, grads[len(feats):]):
        if "gnone_" + n in g:
            assert gr is None or float(gr.abs().max()) == 0.0, n
            continue
        flat = gr.reshape(-1)
        stride = max(1, flat.numel() // 4096)
        # absolute floor: e.g. adapter.*.4.bias is cancelled exactly by the InstanceNorm that follows,
        # so its gradient is pure round-off noise (~1e-9) in both implementations.
        ref = torch.from_numpy(g["gsamp_" + n]).double()
        err = float((flat[::stride].double() - ref).norm())
        assert err <= 2e-3 * float(ref.norm()) + 1e-7 * ref.numel() ** 0.5, n
        assert abs(float(gr.double().norm()) - float(g["gnorm_" + n])) <= 2e-3 * float(g["gnorm_" + n]) + 1e-7 * gr.numel() ** 0.5, n


@pytest.mark.parametrize("name", list(HEAD_CASES))
def test_oracle_fcos_family_head_matches_reference_golden(name):
    """oracle.fcos_head against the unmodified FCOSHead / POTOHead (thirdparty_heads/fcos.py:433-546, poto.py:523-625):
    outputs and every gradient (centerness on either tower, relu*stride and exp box decodings, no centerness branch)."""
    g, (cls, ctr_on_reg, norm_reg, B, hws, strides, _, _), sd, feats, cots = load_head_case(name)
    fx = [f.clone().requires_grad_(True) for f in feats]
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits, bbox_reg, ctr = O.fcos_head(sdo, fx, strides, ctr_on_reg, norm_reg)
    assert (ctr is None) == (cls == "POTOHead")
    groups = [("logits", logits), ("bbox_reg", bbox_reg)] + ([("centerness", ctr)] if ctr is not None else [])
    for gname, outs in groups:
        for l, o in enumerate(outs):
            assert rel_l2(o.detach(), g["%s_%d" % (gname, l)]) < FP32_TOL, (gname, l)
    total = sum((o * c).sum() for (_, outs), cg in zip(groups, cots) for o, c in zip(outs, cg))
    names = sorted(sdo)
    grads = torch.autograd.grad(total, fx + [sdo[n] for n in names])
    for l in range(len(hws)):
        assert rel_l2(grads[l], g["gfeat_%d" % l]) < 1e-4, l
    for n, gr in zip(names, grads[len(fx):]):
        flat = gr.reshape(-1)
        ref = torch.from_numpy(g["gsamp_" + n]).double()
        assert float((flat[::max(1, flat.numel() // 4096)].double() - ref).norm()) <= 1e-4 * float(ref.norm()) + 1e-9, n
        assert abs(float(gr.double().norm()) - float(g["gnorm_" + n])) <= 1e-4 * float(g["gnorm_" + n]) + 1e-9, n


def test_zero_size_boxes_give_empty_masks():
    b = torch.tensor([[32.0, 32.0, 32.0, 96.0], [8.0, 8.0, 24.0, 24.0]])
    m = O.inside_mask(b, (128, 160), (16, 20))
    assert m[0].sum() == 0          # zero width -> inf/NaN distances -> empty (utils.py:87-88)
    assert m[1].sum() == 9          # x,y in {1,2,3}: inclusive on both ends, no half-pixel offset


def test_descriptor_context_row():
    inst = [synth.Instances(torch.tensor([[10.0, 20.0, 50.0, 60.0]]), torch.tensor([3]))]
    (b, oh, lab), = O.prepare_boxes(inst, 800, 1344, add_context_box=True)
    d = O.encode_descriptors(b, oh, 800, 1344)
    assert b[-1].tolist() == [0.0, 0.0, 1343.0, 799.0]
    assert torch.all(d[-1, 4:] == -1) and d[-1, 0] == -1 and d[-1, 1] == -1
    assert abs(float(d[-1, 2]) - 0.99851) < 1e-4 and abs(float(d[-1, 3]) - 0.9975) < 1e-4
    assert lab.tolist() == [3]


def test_norm_classes_descriptor_and_context_box_failure():
    """CATEGORY_FORMAT norm_classes (label_encoder.py:24-25,91-93): one class column, class index / 80; an image without
    GT encodes class 0; with the context box the reference's torch.cat of (N+1, 4) boxes and (N, 1) classes raises."""
    inst = [synth.Instances(torch.tensor([[10.0, 20.0, 50.0, 60.0], [0.0, 0.0, 8.0, 8.0]]), torch.tensor([40, 79])),
            synth.Instances(torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64))]
    per_img = O.prepare_boxes(inst, 800, 1344, add_context_box=False, category_format="norm_classes")
    d = torch.cat([O.encode_descriptors(p[0], p[1], 800, 1344) for p in per_img], 0)
    assert d.shape == (3, 5)
    assert torch.equal(d[:, 4], 2.0 * (torch.tensor([40.0, 79.0, 0.0]) / 80.0 - 0.0) + (-1.0))
    with pytest.raises(RuntimeError):
        O.prepare_boxes(inst, 800, 1344, add_context_box=True, category_format="norm_classes")
    with pytest.raises(ValueError):
        O.prepare_boxes(inst, 800, 1344, add_context_box=False, category_format="bogus")


def test_unknown_pattern_raises():
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case("ctx_stu_adv")
    with pytest.raises(ValueError):
        O.distill_step(sd, bi, im, feats, interact_pattern="bogus")


def test_tf32_operand_rounding_noise_floor():
    """Why the GPU gradient tolerances are what they are: the oracle with TF32-rounded conv
    operands (what the tcgen05 kernels consume) against the same oracle in plain fp32, on a golden case.
    Forward tensors move by a few 1e-4 (inside the 1e-3 bar); gradients of layers below a ReLU move by ~1e-2 because
    that perturbation flips a small fraction of ReLU mask bits -- independent of any CUDA code."""
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case("ctx_stu_adv")

    def run(tf32):
        f = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss, _ = O.distill_step(sd, bi, im, f, distill_flag=flag, tf32=tf32, **cfg_kw)
        cot = synth.synth_cotangents(tea)
        total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
        gf = torch.autograd.grad(total, list(f.values()))
        return tea, float(loss), gf

    tea32, loss32, g32 = run(False)
    teatf, losstf, gtf = run(True)
    assert abs(loss32 - losstf) <= 1e-4 * loss32
    fwd = max(rel_l2(teatf[k].detach(), tea32[k].detach()) for k in tea32)
    grad = max(rel_l2(a, b) for a, b in zip(gtf, g32))
    assert 5e-5 < fwd < 1e-3, fwd          # TF32 forward error: measurable, inside the parity bar
    assert 1e-3 < grad < 8e-2, grad        # ReLU-mask flips amplify it on the gradients (GRAD_TOL_FP32_REFERENCE)


def test_local_inst_conv_over_the_rendered_map_equals_per_box_tap_sums():
    """the input of local_inst_proj_2D is piecewise constant over box rectangles
    (dynamic_teacher.py:137-143), so conv3x3(rendered)[y, x] = b + sum_t sum_tap [(y+dy, x+dx) in box_t] * (W_tap e_t),
    and the gradients follow from ONE box sum with nine accumulators: S[t][tap] = sum_{q in box_t} g[q - tap],
    d e_t = sum_tap W_tap^T S[t][tap], d W_tap = sum_t S[t][tap] (x) e_t. Checked against F.conv2d + autograd in fp64."""
    gen = torch.Generator().manual_seed(12)
    H, W, Cc, T = 9, 13, 6, 5
    boxes = torch.tensor([[0, 0, 13, 9], [2, 1, 7, 5], [12, 8, 13, 9], [5, 3, 6, 9], [3, 0, 11, 1]])   # x0, y0, x1, y1
    e = torch.randn(T, Cc, generator=gen, dtype=torch.float64, requires_grad=True)
    Wt = torch.randn(Cc, Cc, 3, 3, generator=gen, dtype=torch.float64, requires_grad=True)
    bias = torch.randn(Cc, generator=gen, dtype=torch.float64)
    mask = torch.zeros(T, H, W, dtype=torch.float64)
    for t, (x0, y0, x1, y1) in enumerate(boxes.tolist()):
        mask[t, y0:y1, x0:x1] = 1.0
    rendered = torch.einsum("tc,thw->chw", e, mask)[None]
    ref = F.conv2d(rendered, Wt, bias, padding=1)[0]
    g = torch.randn(Cc, H, W, generator=gen, dtype=torch.float64)
    ge_ref, gw_ref = torch.autograd.grad((ref * g).sum(), [e, Wt])
    # forward from per-box tap vectors V[t][tap] = W_tap e_t
    V = torch.einsum("oikl,ti->tklo", Wt.detach(), e.detach())          # (T, 3, 3, C)
    out = bias[:, None, None].expand(Cc, H, W).clone()
    S = torch.zeros(T, 3, 3, Cc, dtype=torch.float64)
    for t, (x0, y0, x1, y1) in enumerate(boxes.tolist()):
        for ky in range(3):
            for kx in range(3):
                dy, dx = ky - 1, kx - 1
                # output pixels p with p + tap inside the box: the box shifted by -tap, clipped to the image
                ya, yb, xa, xb = max(y0 - dy, 0), min(y1 - dy, H), max(x0 - dx, 0), min(x1 - dx, W)
                if yb > ya and xb > xa:
                    out[:, ya:yb, xa:xb] += V[t, ky, kx][:, None, None]
                    S[t, ky, kx] = g[:, ya:yb, xa:xb].sum((1, 2))      # = sum over q in the box of g[q - tap]
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    ge = torch.einsum("tklo,oikl->ti", S, Wt.detach())
    gw = torch.einsum("tklo,ti->oikl", S, e.detach())
    assert torch.allclose(ge, ge_ref, rtol=1e-12, atol=1e-12)
    assert torch.allclose(gw, gw_ref, rtol=1e-12, atol=1e-12)


# ----------------------------------------------------------------------------- tap rendering: coverage rule
# Python mirror of tap_paint_kernel's mask construction (csrc/taprender.cu): per strip and row an `any` mask (box dilated by
# one pixel), an `all` mask (eroded by one pixel) and the marks of the pixels whose set of in-box taps can differ from the
# left neighbour's; a thread evaluates its first pixel and marked pixels and copies otherwise.
def bit_run(n): return 0xffffffff if n >= 32 else (1 << n) - 1
def strip_masks(r, W, HW, pix0):
    """mirror of tap_paint_kernel's per-row mask construction for one strip; r = (x0, x1, y0, y1) half-open"""
    x0, x1, y0, y1 = r
    ya, xa = divmod(pix0, W)
    npx = min(32, HW - pix0)
    m_any = m_all = mark = 0
    box = x1 > x0 and y1 > y0
    x, y, j0 = xa, ya, 0
    while j0 < npx:
        run = min(W - x, npx - j0)
        mark |= 1 << j0
        if box and y >= y0 - 1 and y < y1 + 1:
            lo, hi = max(x0 - 1, x), min(x1 + 1, x + run)
            if hi > lo:
                m_any |= (bit_run(hi - lo) << (j0 + lo - x)) & 0xffffffff
                for pos in (x0 - 1, x0, x0 + 1, x1 - 1, x1, x1 + 1):
                    if x <= pos < x + run:
                        mark |= 1 << (j0 + pos - x)
            if y >= y0 + 1 and y < y1 - 1:
                lo2, hi2 = max(x0 + 1, x), min(x1 - 1, x + run)
                if hi2 > lo2:
                    m_all |= (bit_run(hi2 - lo2) << (j0 + lo2 - x)) & 0xffffffff
        j0 += run; x = 0; y += 1
    return m_any, m_all, mark, npx

def tapset_true(r, y, x):
    x0, x1, y0, y1 = r
    return frozenset((dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if y0 <= y + dy < y1 and x0 <= x + dx < x1)

def kernel_tapset(r, m_any, m_all, j, y, x):
    if not (m_any >> j) & 1: return frozenset()
    if (m_all >> j) & 1: return frozenset((dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    return tapset_true(r, y, x)   # ring: per-pixel evaluation, as in the kernel



def test_tap_render_coverage_masks_and_copy_rule_reproduce_every_pixels_tap_set():
    """Brute force over random boxes (degenerate, full image, borders) on narrow and wide levels: the tap set the kernel
    would use for every pixel -- from the masks, the ring evaluation and the copy-from-the-left rule -- equals the set
    {(dy, dx): pixel + (dy, dx) inside the box}."""
    import random
    rng = random.Random(1)
    for trial in range(1200):
        H, W = rng.choice([(1, 2), (2, 3), (4, 5), (7, 11), (13, 21), (25, 42), (9, 40), (3, 33), (5, 32), (6, 31)])
        HW = H * W
        rows = []
        for _ in range(rng.randint(0, 6)):
            xa, xb = sorted(rng.randint(0, W) for _ in range(2))
            ya, yb = sorted(rng.randint(0, H) for _ in range(2))
            if rng.random() < 0.2:
                xb = xa
            if rng.random() < 0.2:
                xa, xb, ya, yb = 0, W, 0, H
            rows.append((xa, xb, ya, yb))
        for pix0 in range(0, HW, 32):
            ms = [strip_masks(r, W, HW, pix0) for r in rows]
            npx = min(32, HW - pix0)
            differs, x, j0 = 0, pix0 % W, 0
            while j0 < npx:               # every thread marks the first pixel of each image row of the strip
                differs |= 1 << j0
                j0 += min(W - x, npx - j0)
                x = 0
            for (_, _, mk, _) in ms:
                differs |= mk
            for sub in range(4):
                cur = None
                for i in range(8):
                    j = sub * 8 + i
                    if j >= npx:
                        break
                    y, x = divmod(pix0 + j, W)
                    if i == 0 or (differs >> j) & 1:
                        cur = tuple(kernel_tapset(r, m[0], m[1], j, y, x) for r, m in zip(rows, ms))
                    assert cur == tuple(tapset_true(r, y, x) for r in rows), ((H, W), rows, pix0, j)
