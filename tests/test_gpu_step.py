"""GPU parity of the whole distillation step through the plugin classes (DynamicTeacher, SequentialConvs,
BaseDistillator.distill_loss) against (1) the golden vectors produced by the unmodified reference and (2) the CPU
oracle on the same seeded inputs.

Parity metric (SURVEY.md 8(d) "parity gate"; north_star: 1e-3 relative fp32, bit-exact label->region assignment):
  * masks / labels: exact;
  * every forward float tensor: ||d||_2/||ref||_2 <= 1e-3 vs the fp32 reference; scalar loss: relative <= 1e-3;
  * gradients: <= 2e-2 relative-L2 vs the oracle run with 10-bit-mantissa (TF32-rounded) conv operands -- the same
    mantissa as the fp16 operands the tcgen05 kernels consume (measured 5.5e-3 on feature gradients, 1.1e-2 on the
    label-side biases that sum few one-pixel boxes: accumulation-order-sized forward differences still flip a ~3e-5
    fraction of ReLU mask bits, and a flipped bit changes its gradient entry by 100 %; each backward kernel in
    isolation is held to 2e-5 in test_gpu_kernels.py).
    Against the un-rounded fp32 reference, gradients of layers that sit below a ReLU differ by ~2e-2 for ANY
    perturbed forward (each flipped ReLU mask bit changes its gradient entry by 100 %); the oracle's operand-rounding
    emulation reproduces that number on the CPU (see DESIGN.md "Precision"), so it is asserted here as a loose bound
    too. engine.FORWARD_PRECISION = "tf32x3" (split-operand, fp32-accurate) is tested separately below, and so is the
    exact linearity of the backward in the incoming gradient scale (fp16 gradient operands carry a power-of-two scale).
"""
import numpy as np
import pytest
import torch

from lgd_b200 import engine, synth
from oracle import lgd_oracle as O
from oracle.make_golden import CASES
from tests.golden_util import load_case, rel_l2, unpack_mask
from tests.gpu_util import run_engine

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-3
GRAD_TOL_TF32_ORACLE = 2e-2
GRAD_TOL_FP32_REFERENCE = 8e-2


@pytest.mark.parametrize("name", list(CASES))
def test_step_matches_reference_golden(name):
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case(name)
    out = run_engine(cfg_kw, sd, bi, im, feats, flag)
    assert abs(out["loss"] - float(g["loss"])) <= FWD_TOL * abs(float(g["loss"]))
    S = out["model"].teacher._last
    assert rel_l2(S.label_embed.cpu(), g["label_embed"]) < 1e-4
    assert rel_l2(S.canoni.cpu(), g["canoni"]) < 1e-4
    for l, k in enumerate(feats):
        got = torch.cat(out["masks"][l], 0)
        assert torch.equal(got, unpack_mask(g, k)), "label->region assignment must be bit exact"
        assert rel_l2(out["tea"][k], g[f"tea_{k}"]) < FWD_TOL, (k, rel_l2(out["tea"][k], g[f"tea_{k}"]))
        T = S.tb.T
        if cfg_kw.get("interact_pattern", "stuGuided") == "stuGuided":
            assert rel_l2(S.pooled.view(-1, T, 256)[l].cpu(), g[f"mha_q_{k}"]) < FWD_TOL
        assert rel_l2(S.a.view(-1, T, 256)[l].cpu(), g[f"mha_out_{k}"]) < FWD_TOL
    for i, il in enumerate(out["inst_labels"]):
        assert np.array_equal(il.cpu().numpy().astype(np.float32), g[f"inst_labels_{i}"])
    # gradients vs the fp32 reference: loose bound (ReLU mask flips, see module docstring)
    for k in feats:
        ref = g[f"gfeat_{k}"]
        got = out["gfeat"][k]
        if ref.size == 0:
            assert got is None or float(got.abs().max()) == 0.0
        else:
            assert rel_l2(got, ref) < GRAD_TOL_FP32_REFERENCE, (k, rel_l2(got, ref))
    samp_err, samp_n = 0.0, 0
    for n, gr in out["gparam"].items():
        if "gnone_" + n in g:
            assert gr is None or float(gr.abs().max()) == 0.0, n
            continue
        assert gr is not None, n
        ref_norm = float(g["gnorm_" + n])
        assert abs(float(gr.double().norm()) - ref_norm) <= GRAD_TOL_FP32_REFERENCE * ref_norm + 1e-7 * gr.numel() ** 0.5, n
        # element samples of the reference's gradient (every stride-th entry, 4096 per tensor), not only its norm.
        # Per tensor only a gross bound can hold against the PLAIN reference on these small inputs: a handful of flipped
        # ReLU decisions on a 2 x 3 pixel level moves a 256-entry bias gradient by 10-20 % (measured 18 % on
        # student_proj_2D.bias of ctx_label_detach, whose gradient against the oracle evaluated with the engine's
        # activation pattern is within 3e-3: tests/test_gpu_parity.py::test_gradient_parity_on_golden_inputs). A layout
        # or indexing error shows as >= 100 %. The flip bound is asserted on all samples of the step together.
        flat = gr.reshape(-1)
        stride = max(1, flat.numel() // 4096)
        ref = torch.from_numpy(g["gsamp_" + n]).double()
        d = flat[::stride].double() - ref
        assert float(d.norm()) <= 0.3 * float(ref.norm()) + 1e-7 * ref.numel() ** 0.5, (n, float(d.norm()), float(ref.norm()))
        if not n.endswith("adapter.4.bias"):   # analytically zero: round-off in both implementations
            samp_err += float(d.pow(2).sum() / ref.pow(2).sum().clamp_min(1e-60))
            samp_n += 1
    assert (samp_err / max(samp_n, 1)) ** 0.5 <= GRAD_TOL_FP32_REFERENCE, (samp_err / max(samp_n, 1)) ** 0.5


@pytest.mark.parametrize("name", ["ctx_stu_adv", "noctx_stu_empty"])
def test_step_gradients_match_tf32_oracle(name):
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case(name)
    out = run_engine(cfg_kw, sd, bi, im, feats, flag)
    f = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    # the emulation rounds the operands of every convolution the engine runs on the tensor cores: all eight, or seven
    # with tap rendering (local_inst_proj_2D evaluated exactly from per-box tap vectors)
    tea, _, _, loss, _ = O.distill_step(sdo, bi, im, f, distill_flag=flag, tf32=True,
                                        exact_local_inst=engine.TAP_RENDER, **cfg_kw)
    cot = synth.synth_cotangents(tea)
    total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
    names = sorted(sdo)
    grads = torch.autograd.grad(total, list(f.values()) + [sdo[n] for n in names], allow_unused=True)
    assert abs(out["loss"] - float(loss)) <= 1e-4 * float(loss)
    worst = 0.0
    for k in tea:
        e = rel_l2(out["tea"][k], tea[k])
        worst = max(worst, e)
        # same rounded operands: fp32 accumulation order differs, and an accumulation-order-sized change flips the
        # TF32 rounding (2^-11 relative) of a few intermediate activations between the chained convolutions
        assert e < 5e-4, (k, e)
    for l, k in enumerate(f):
        if grads[l] is not None:
            e = rel_l2(out["gfeat"][k], grads[l])
            assert e < GRAD_TOL_TF32_ORACLE, (k, e)
    for n, gr in zip(names, grads[len(f):]):
        got = out["gparam"][n]
        if gr is None:
            assert got is None or float(got.abs().max()) == 0.0, n
            continue
        err = float((got.double() - gr.double()).norm())
        assert err <= GRAD_TOL_TF32_ORACLE * float(gr.double().norm()) + 1e-7 * gr.numel() ** 0.5, (n, err, float(gr.norm()))


@pytest.mark.parametrize("name", list(CASES))
def test_step_tf32x3_mode_tightens_every_tensor(name, monkeypatch):
    """engine.FORWARD_PRECISION = "tf32x3": all convolutions of the forward and the dgrad chain run split-operand TF32
    (three chained launches, ~4e-6 per convolution). Against the plain fp32 oracle (which reproduces the reference's
    goldens, tests/test_oracle.py) the teacher pyramid then agrees to 2e-5 and the loss to 1e-5 (default mode: 5e-4 /
    1e-4) -- asserted. Gradients (tools/grad_err_probe.py prints them per parameter): where no ReLU mask bit differs
    from the oracle's they drop to the TF32 rounding of the wgrad operands (3e-4 on conv weights) or below (5e-5 on
    the feature gradients and < 3e-4 on the whole label side in the cases without a flip); what remains elsewhere is
    the discrete flip floor of these small cases -- ONE flipped bit among the ~2e5 activations of a layer is
    sqrt(1/2e5) = 2.2e-3 and more when it sits on a heavy-tailed gradient entry (measured: <= 2.8e-3 / 1.3e-2 /
    5.7e-3 on the three cases, default mode 1e-2 .. 4e-2) -- so the gradient bound asserted is the default mode's.
    The backward kernels themselves are held to 2e-5 in test_gpu_kernels.py."""
    from lgd_b200 import engine
    monkeypatch.setattr(engine, "FORWARD_PRECISION", "tf32x3")
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case(name)
    out = run_engine(cfg_kw, sd, bi, im, feats, flag)
    f = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tea, _, _, loss, _ = O.distill_step(sdo, bi, im, f, distill_flag=flag, tf32=False, **cfg_kw)
    cot = synth.synth_cotangents(tea)
    total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
    names = sorted(sdo)
    grads = torch.autograd.grad(total, list(f.values()) + [sdo[n] for n in names], allow_unused=True)
    assert abs(out["loss"] - float(loss)) <= 1e-5 * float(loss)
    assert abs(out["loss"] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k in tea:
        assert rel_l2(out["tea"][k], tea[k]) < 2e-5, (k, rel_l2(out["tea"][k], tea[k]))
        assert rel_l2(out["tea"][k], g[f"tea_{k}"]) < 2e-5
    worst, fails = {}, []
    for l, k in enumerate(f):
        if grads[l] is None:
            assert out["gfeat"][k] is None or float(out["gfeat"][k].abs().max()) == 0.0
            continue
        e = rel_l2(out["gfeat"][k], grads[l])
        worst["gfeat_" + k] = e
        assert e < GRAD_TOL_TF32_ORACLE, (k, e)
    for n, gr in zip(names, grads[len(f):]):
        got = out["gparam"][n]
        if gr is None:
            assert got is None or float(got.abs().max()) == 0.0, n
            continue
        err = float((got.double() - gr.double()).norm())
        ref = float(gr.double().norm())
        worst[n] = err / max(ref, 1e-30)
        # the absolute term only matters for adapter.4.bias, whose gradient is analytically zero (round-off in both)
        if err > GRAD_TOL_TF32_ORACLE * ref + 1e-7 * gr.numel() ** 0.5:
            fails.append((n, err, ref))
    print("tf32x3 worst relative gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert not fails, fails


@pytest.mark.parametrize("alpha,tol", [(2.0 ** -27, 1e-6), (2.0 ** 20, 1e-6), (1e-8, 2e-3), (1e6, 2e-3)])
def test_backward_is_linear_over_fourteen_decades(alpha, tol):
    """Size-independent property of the backward: gradients are linear in the incoming gradient. The convolutions of
    the backward run on fp16 copies of the gradient tensors, scaled by a power of two chosen from a norm bound -- so
    scaling the loss gradient and the teacher cotangents by 1e-8 or 1e6 (far outside fp16's range either way) must
    scale every parameter and feature gradient by that factor: exactly for a power of two (every fp16 conversion then
    sees the same mantissas and only the power-of-two scale moves), and within the 10-bit operand rounding otherwise."""
    from tests.gpu_util import make_model
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case("ctx_stu_adv")

    def grads(a):
        m = make_model(cfg_kw, sd, flag)
        f = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss = m.forward(bi, im, f)
        cot = synth.synth_cotangents({k: v.detach().cpu() for k, v in tea.items()})
        keys = list(tea.keys())
        torch.autograd.backward([loss] + [tea[k] for k in keys],
                                [torch.full_like(loss, a)] + [(cot[k] * a).cuda() for k in keys])
        torch.cuda.synchronize()
        out = {n: p.grad.double().cpu() for n, p in m.named_parameters() if p.grad is not None}
        out.update({"feat_" + k: v.grad.double().cpu() for k, v in f.items()})
        return out

    base, scaled = grads(1.0), grads(alpha)
    assert base.keys() == scaled.keys()
    for n in base:
        assert torch.isfinite(scaled[n]).all(), n
        if n.endswith("adapter.4.bias") and tol > 1e-5:
            continue   # analytically zero (InstanceNorm removes channel constants): pure round-off, not linear in alpha
        ref = base[n] * alpha
        err = float((scaled[n] - ref).norm())
        assert err <= tol * float(ref.norm()) + 1e-30, (n, err, float(ref.norm()))


def test_plugin_surface_and_eval_mode():
    """Registry names resolve, state_dict names/shapes are the reference's, forward works under no_grad, an image
    without GT takes the dummy-box path, unknown patterns raise ValueError."""
    import lgd_b200
    cfg = synth.make_cfg(device="cuda")
    t = lgd_b200.CUSTOMIZED_DETECTORS_REGISTRY.get("DynamicTeacher")(cfg).cuda().eval()
    a = lgd_b200.build_adapter(cfg)
    shapes = synth.hot_path_param_shapes()
    for k, v in t.state_dict().items():
        assert tuple(v.shape) == shapes["teacher." + k]
    for k, v in a.state_dict().items():
        assert tuple(v.shape) == shapes["adapter.distill." + k]
    bi, im, feats = synth.synth_batch(2, 100, 130, seed=9, n_boxes=[0, 3], feature_device="cuda")
    with torch.no_grad():
        tea, labels, masks = t((bi, im, None, feats))
    assert list(tea.keys()) == list(feats.keys())
    for k in feats:
        assert tea[k].shape == feats[k].shape and bool(torch.isfinite(tea[k]).all())
    assert len(masks) == 5 and len(masks[0]) == 2 and masks[0][0].shape[0] == 1 and masks[0][1].shape[0] == 4
    assert labels[0].tolist() == [0.0] and labels[1].shape[0] == 3
    # stand-alone adapter call (hook API) == first level of the fused path
    y = a(feats["p5"])
    assert y.shape == feats["p5"].shape
    t.interact_pattern = "bogus"
    with pytest.raises(ValueError):
        t((bi, im, None, feats))


def test_norm_classes_with_context_box_raises_like_the_reference():
    """label_encoder.py:75-77,91-93,105: CATEGORY_FORMAT norm_classes + ADD_CONTEXT_BOX concatenates (N+1, 4) boxes with
    (N, 1) classes -- the reference raises RuntimeError for every image with GT; so does the engine (the working
    combination, without the context box, is the golden case noctx_stu_normcls)."""
    sd = synth.synth_state_dict(5, desc_dim=5)
    bi, im, feats = synth.synth_batch(2, 96, 128, seed=3, n_boxes=[2, 1])
    with pytest.raises(RuntimeError):
        run_engine(dict(add_context_box=True, category_format="norm_classes"), sd, bi, im, feats, 1, backward=False)


def test_full_size_properties():
    """BASELINE config sizes (800x1333, B=2): size-independent properties instead of a CPU oracle run:
    teacher pyramid is GroupNorm-normalised per image and level (mean 0, var 1), loss is finite and invariant to
    a uniform scaling of the teacher's last conv (GN) -- and a second identical step is bit-identical."""
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(2, 800, 1333, seed=1234)
    a = run_engine(dict(add_context_box=True), sd, bi, im, feats, 1, backward=True)
    b = run_engine(dict(add_context_box=True), sd, bi, im, feats, 1, backward=True)
    assert a["loss"] == b["loss"]
    for k in a["tea"]:
        assert torch.equal(a["tea"][k], b["tea"][k]), "forward must be deterministic"
        assert torch.equal(a["gfeat"][k], b["gfeat"][k]), "backward must be deterministic"
        t = a["tea"][k].double().flatten(1)
        assert float(t.mean(1).abs().max()) < 1e-4
        assert float((t.var(1, unbiased=False) - 1).abs().max()) < 1e-3
    assert np.isfinite(a["loss"]) and 0.5 < a["loss"] < 4.0


@pytest.mark.parametrize("cfg_kw,hw,B", [
    (dict(add_context_box=True), (800, 1333), 1),      # BASELINE configs[0]/[1] image size (RetinaNet, ctx box)
    (dict(add_context_box=False), (640, 1067), 2),     # configs[2] FCOS (no ctx box) at configs[4]'s smallest scale
])
def test_forward_matches_oracle_at_baseline_sizes(cfg_kw, hw, B):
    """Full-resolution forward parity against the CPU oracle (fp32 reference arithmetic) -- sizes the oracle finishes
    in seconds: masks bit exact, teacher pyramid and loss within the 1e-3 bar."""
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(B, hw[0], hw[1], seed=77)
    out = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=False)
    with torch.no_grad():
        tea_o, _, masks_o, loss_o, _ = O.distill_step(sd, bi, im, feats, **cfg_kw)
    assert abs(out["loss"] - float(loss_o)) <= FWD_TOL * float(loss_o)
    for l, k in enumerate(feats):
        assert torch.equal(torch.cat(out["masks"][l], 0), torch.cat(masks_o[l], 0))
        assert rel_l2(out["tea"][k], tea_o[k]) < FWD_TOL, (k, rel_l2(out["tea"][k], tea_o[k]))


def test_batch_and_shape_changes_between_steps():
    """Dynamic shapes (multi-scale training, varying box counts): consecutive steps with different batch sizes, image
    sizes and T reuse the same module; each must equal a fresh run of that step alone (no stale per-shape state)."""
    sd = synth.synth_state_dict(5)
    cfg_kw = dict(add_context_box=True)
    from tests.gpu_util import make_model
    m = make_model(cfg_kw, sd, 1)
    cases = [synth.synth_batch(2, 200, 264, seed=1), synth.synth_batch(3, 160, 232, seed=2, n_boxes=[0, 1, 40]),
             synth.synth_batch(1, 96, 96, seed=3), synth.synth_batch(2, 200, 264, seed=4)]
    for bi, im, feats in cases:
        f = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss = m.forward(bi, im, f)
        loss.backward()
        ref = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=False)
        assert float(loss) == ref["loss"]
        for k in tea:
            assert torch.equal(tea[k].detach().cpu(), ref["tea"][k])
        m.zero_grad(set_to_none=True)


def test_many_boxes_per_image():
    """Crowded images (64 boxes + context box per image, the synthetic workload's cap): T = 260 tokens, long key lists
    in the block-diagonal attention, many overlapping boxes per pixel in rendering. Forward parity vs the oracle and a
    finite, deterministic backward."""
    sd = synth.synth_state_dict(5)
    cfg_kw = dict(add_context_box=True)
    bi, im, feats = synth.synth_batch(4, 160, 200, seed=21, n_boxes=[64, 64, 64, 64])
    a = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=True)
    b = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=True)
    with torch.no_grad():
        tea_o, _, masks_o, loss_o, _ = O.distill_step(sd, bi, im, feats, **cfg_kw)
    assert abs(a["loss"] - float(loss_o)) <= FWD_TOL * float(loss_o)
    for l, k in enumerate(feats):
        assert torch.equal(torch.cat(a["masks"][l], 0), torch.cat(masks_o[l], 0))
        assert rel_l2(a["tea"][k], tea_o[k]) < FWD_TOL, (k, rel_l2(a["tea"][k], tea_o[k]))
        assert torch.equal(a["gfeat"][k], b["gfeat"][k])
        assert bool(torch.isfinite(a["gfeat"][k]).all())
    for n, gr in a["gparam"].items():
        assert gr is not None and bool(torch.isfinite(gr).all()), n
        assert torch.equal(gr, b["gparam"][n]), n


@pytest.mark.parametrize("pattern", ["student_fill", "teacher_fill"])
def test_fill_patterns_match_oracle(pattern):
    """INTERACT_PATTERN student_fill / teacher_fill (dynamic_teacher.py:261-264) bypass the attention block."""
    sd = synth.synth_state_dict(5)
    cfg_kw = dict(add_context_box=True, interact_pattern=pattern)
    bi, im, feats = synth.synth_batch(2, 120, 150, seed=31, n_boxes=[3, 7])
    out = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=True)
    with torch.no_grad():
        tea_o, _, _, loss_o, _ = O.distill_step(sd, bi, im, feats, **cfg_kw)
    assert abs(out["loss"] - float(loss_o)) <= FWD_TOL * float(loss_o)
    for k in feats:
        assert rel_l2(out["tea"][k], tea_o[k]) < FWD_TOL, (k, rel_l2(out["tea"][k], tea_o[k]))
    # student_fill never reads the label embeddings, teacher_fill never reads the pooled appearance embeddings
    unused = ("multi_head_attn",) + (("label_encoder_", "canoni_proj") if pattern == "student_fill" else ("student_proj_2D",))
    for n, gr in out["gparam"].items():
        if any(u in n for u in unused):
            assert gr is None or float(gr.abs().max()) == 0.0, n      # DDP find_unused_parameters contract
        else:
            assert gr is not None and bool(torch.isfinite(gr).all()), n


def test_rcnn_style_pyramid_p2_to_p6():
    """Faster R-CNN style pyramid: five levels starting at stride 4, detached appearance
    embeddings (configs/Distillation/FasterRCNN): forward parity vs the oracle, no gradient into the student maps from
    the teacher branch."""
    sd = synth.synth_state_dict(5)
    cfg_kw = dict(add_context_box=True, detach_appearance_embed=True)
    bi, im, _ = synth.synth_batch(2, 128, 160, seed=41)
    gen = torch.Generator().manual_seed(42)
    feats = {k: torch.randn(2, 256, h, w, generator=gen) for k, (h, w) in
             zip(("p2", "p3", "p4", "p5", "p6"), [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)])}
    out = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=True)
    with torch.no_grad():
        tea_o, _, masks_o, loss_o, _ = O.distill_step(sd, bi, im, feats, **cfg_kw)
    assert abs(out["loss"] - float(loss_o)) <= FWD_TOL * float(loss_o)
    for l, k in enumerate(feats):
        assert torch.equal(torch.cat(out["masks"][l], 0), torch.cat(masks_o[l], 0))
        assert rel_l2(out["tea"][k], tea_o[k]) < FWD_TOL, (k, rel_l2(out["tea"][k], tea_o[k]))
        assert out["gfeat"][k] is not None      # the distillation loss still reaches the student maps
    assert out["gparam"]["teacher.student_proj_2D.0.0.weight"] is not None


def test_rcnn_pyramid_at_full_size():
    """The same at the published size (Faster / Mask R-CNN recipes: 800x1344, P2-P6 = 200x336 ... 13x21, 89 523 pixels per
    image): the stride-4 level is wider than two convolution tiles (three-run input strips, 529 tiles per image). Forward
    parity vs the oracle, a complete backward."""
    sd = synth.synth_state_dict(5)
    cfg_kw = dict(add_context_box=True, detach_appearance_embed=True)
    bi, im, _ = synth.synth_batch(1, 800, 1333, seed=43)
    gen = torch.Generator().manual_seed(44)
    hws = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    feats = {k: torch.randn(1, 256, h, w, generator=gen) for k, (h, w) in zip(("p2", "p3", "p4", "p5", "p6"), hws)}
    out = run_engine(cfg_kw, sd, bi, im, feats, 1, backward=True)
    with torch.no_grad():
        tea_o, _, masks_o, loss_o, _ = O.distill_step(sd, bi, im, feats, **cfg_kw)
    assert abs(out["loss"] - float(loss_o)) <= FWD_TOL * float(loss_o)
    for l, k in enumerate(feats):
        assert torch.equal(torch.cat(out["masks"][l], 0), torch.cat(masks_o[l], 0))
        assert rel_l2(out["tea"][k], tea_o[k]) < FWD_TOL, (k, rel_l2(out["tea"][k], tea_o[k]))
        assert out["gfeat"][k] is not None and bool(torch.isfinite(out["gfeat"][k]).all())
    for n, gr in out["gparam"].items():
        assert gr is None or bool(torch.isfinite(gr).all()), n
