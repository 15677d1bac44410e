"""Gradient parity of the whole step at BASELINE sizes (800x1333 -> 800x1344, B = 2; RetinaNet with the context box
and the FCOS variant without it), through the plugin classes and the C ABI, against the CPU oracle run on the box's
host cores (~3 s per oracle step).

Bars (oracle/parity.py explains the method; DESIGN.md section 6 the numbers):
  * label->region assignment bit exact; loss and teacher pyramid within 1e-3 of the fp32 oracle;
  * FLIP FRACTION: share of the 6.9e7 ReLU decisions of the step that differ from the fp32 oracle's <= 2e-4
    (measured 6e-5 .. 9e-5 on fp16 operands, 5e-7 in tf32x3 mode), and every flipped entry is a rounding-level tie:
    the oracle's own pre-activation there is <= 5e-3 of the tensor's rms (measured <= 2.5e-3; tf32x3: <= 1e-5);
  * KERNEL ACCURACY: every feature and parameter gradient against the oracle evaluated with the engine's activation
    pattern: <= 1e-3 in tf32x3 mode (measured <= 3.1e-4: what is left is the single-pass TF32 rounding of the wgrad
    operands), <= 2e-3 on the default fp16 operands (measured 8e-4 .. 1.5e-3 at the end of the longest chains, the feature
    gradients, adapter.0.weight and the label encoder's first layer: 10-bit-mantissa operand rounding, ~4e-4 per convolution, accumulated in quadrature
    along the 5 forward + 5 backward convolutions of the teacher chain -- the same mantissa the stock reference computes
    with on any Ampere+ GPU, where cuDNN runs its convolutions in TF32);
  * against the PLAIN fp32 oracle (flips included) the global figures: <= 8e-2 on fp16 operands (measured 3e-2 .. 6e-2),
    <= 5e-3 in tf32x3 mode (measured 9e-4 .. 2.6e-3 from ~30 flipped decisions out of 6.9e7)."""
import pytest

from oracle import parity

pytestmark = pytest.mark.gpu

CASES = {"retinanet_ctx": dict(add_context_box=True), "fcos_noctx": dict(add_context_box=False)}


def _report(tag, r):
    print("%s: loss %.1e fwd %.1e flips %d/%d = %.1e (margin %.1e) grad|pattern %.2e (%s) grad|plain %.2e (%s)"
          % (tag, r["loss_err"], r["fwd_err"], r["flips"], r["activations"], r["flip_fraction"], r["flip_margin"],
             r["grad_err_pattern"], r["grad_err_pattern_worst"], r["grad_err_plain"], r["grad_err_plain_worst"]))


@pytest.mark.parametrize("name", list(CASES))
def test_gradient_parity_at_baseline_size_fp16_operands(name, monkeypatch):
    from lgd_b200 import engine
    monkeypatch.setattr(engine, "FORWARD_PRECISION", "fp16")
    r = parity.step_parity(CASES[name], 2, (800, 1333), seed=77)
    _report("fp16 " + name, r)
    assert r["masks_exact"]
    assert r["loss_err"] <= 1e-3 and r["fwd_err"] <= 1e-3
    assert r["activations"] == 6 * 2 * 256 * 22400
    assert r["flip_fraction"] <= 2e-4, r["flips_per_site"]
    assert r["flip_margin"] <= 5e-3
    assert r["grad_err_pattern"] <= 2e-3, sorted(r["table_pattern"].items(), key=lambda kv: -kv[1])[:5]
    # everything that does not sit at the far end of a backward chain meets the 1e-3 bar outright: above it (and below
    # 2e-3) are only the feature gradients and adapter.0.weight (five forward + five backward convolutions away) and the
    # first layer of the label encoder (the five teacher convolutions, the relation block and the whole encoder stack
    # away; measured 1.05e-3 / 1.00e-3 on its bias / weight in the case without the context box)
    far_end = ("feat/", "adapter.distill.adapter.0.weight", "teacher.label_encoder_.conv1.",
               "teacher.label_encoder_.stn_desc.conv1.")
    over = sorted((k, round(v, 6)) for k, v in r["table_pattern"].items() if v > 1e-3)
    assert all(k.startswith(far_end) for k, _ in over), str(over)
    assert r["grad_err_plain"] <= 8e-2


@pytest.mark.parametrize("name", list(CASES))
def test_gradient_parity_at_baseline_size_tf32x3(name, monkeypatch):
    from lgd_b200 import engine
    monkeypatch.setattr(engine, "FORWARD_PRECISION", "tf32x3")
    r = parity.step_parity(CASES[name], 2, (800, 1333), seed=77)
    _report("tf32x3 " + name, r)
    assert r["masks_exact"]
    assert r["loss_err"] <= 1e-5 and r["fwd_err"] <= 2e-5
    assert r["flip_fraction"] <= 2e-6, r["flips_per_site"]
    assert r["flip_margin"] <= 1e-4
    assert r["grad_err_pattern"] <= 1e-3, sorted(r["table_pattern"].items(), key=lambda kv: -kv[1])[:5]
    assert r["grad_err_plain"] <= 5e-3


@pytest.mark.parametrize("name", ["ctx_stu_adv", "noctx_stu_empty", "ctx_label_detach", "ctx_stu_x1y1wh", "seg_ctx_detach",
                                  "noctx_stu_normcls"])
def test_gradient_parity_on_golden_inputs(name):
    """The same method on the inputs of the reference goldens (adversarial boxes, an image without GT, labelGuided with
    detached appearance embeddings and distill_flag = 0): small tensors, so a single flipped decision is already
    ~2e-3 of a layer -- only the pattern-evaluated comparison is meaningful here."""
    from tests.golden_util import load_case
    g, cfg_kw, batch_kw, flag, sd, bi, im, feats = load_case(name)
    r = parity.step_parity_inputs(cfg_kw, sd, bi, im, feats, flag)
    _report("golden " + name, r)
    assert r["masks_exact"]
    assert r["loss_err"] <= 1e-3 and r["fwd_err"] <= 1e-3
    assert r["flip_fraction"] <= 5e-4 and r["flip_margin"] <= 1e-2
    # Levels of a handful of pixels (P7 / P6 of these 100..128-pixel images are 1x2 / 2x3): InstanceNorm over two pixels maps
    # every channel to +-1/sqrt(1 + 4 eps/(a-b)^2), so the loss gradient there comes from the few channels with a ~ b and
    # is ill-conditioned in the forward values themselves (any 1e-4 change of the teacher pyramid moves it by percents).
    # Measured on the convolution path: 6.8e-3 (P7) / 2.0e-3 (P6) on noctx_stu_normcls, <= 3e-3 on the others. Held to 2e-2
    # -- a gross-error bar for an ill-conditioned quantity; every other tensor to 3e-3.
    tiny = {"feat/" + k for k, v in feats.items() if v.shape[-2] * v.shape[-1] < 8}
    table = sorted(r["table_pattern"].items(), key=lambda kv: -kv[1])
    assert all(v <= (2e-2 if k in tiny else 3e-3) for k, v in table), table[:5]


def test_forward_operand_range_large_inputs_keep_parity_and_overflow_is_loud():
    """The forward convolutions read fp16 copies of their inputs (5-bit exponent, max 65504). FPN maps a thousand
    times larger than usual are still far inside that range everywhere on the path (GroupNorm / InstanceNorm make the
    teacher and the loss scale free), so parity must hold unchanged. Inputs that DO leave the range become inf in the
    fp16 copy; every ReLU of the path lets NaN through (relu_keep_nan) and every statistic is computed from the
    affected values, so the loss and the teacher pyramid come out non-finite -- the training loop aborts on that
    (train.py:194) -- never as a finite, silently wrong number."""
    import torch
    from lgd_b200 import synth
    from oracle import lgd_oracle as O
    from tests.gpu_util import rel_l2, run_engine
    sd = synth.synth_state_dict(5)
    cfg_kw = dict(add_context_box=True)
    bi, im, feats = synth.synth_batch(2, 160, 200, seed=55)
    big = {k: v * 1e3 for k, v in feats.items()}
    out = run_engine(cfg_kw, sd, bi, im, big, 1, backward=True)
    with torch.no_grad():
        tea_o, _, _, loss_o, _ = O.distill_step(sd, bi, im, big, **cfg_kw)
    assert abs(out["loss"] - float(loss_o)) <= 1e-3 * float(loss_o)
    for k in big:
        assert rel_l2(out["tea"][k], tea_o[k]) < 1e-3, (k, rel_l2(out["tea"][k], tea_o[k]))
        assert bool(torch.isfinite(out["gfeat"][k]).all())
    huge = {k: v * 1e6 for k, v in feats.items()}
    bad = run_engine(cfg_kw, sd, bi, im, huge, 1, backward=False)
    assert not torch.isfinite(torch.tensor(bad["loss"])), "an overflowed forward must not produce a finite loss"
    for k in huge:
        assert not bool(torch.isfinite(bad["tea"][k]).all()), k
