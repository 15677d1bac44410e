"""The native step runtime (lgd_b200/csrc/chain.cu: one C-ABI call per chain) against the per-kernel orchestration in
lgd_b200/engine.py: same kernels in the same order, so every output and every gradient must be BIT-IDENTICAL -- with
and without a context box, with detached appearance embeddings, with distill_flag = 0, with an image without GT."""
import pytest
import torch

from lgd_b200 import engine, synth
from tests.gpu_util import run_engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg_kw,batch_kw,flag", [
    (dict(add_context_box=True), dict(B=2, img_h=120, img_w=150, seed=101, adversarial=True), 1),
    (dict(add_context_box=False), dict(B=3, img_h=100, img_w=130, seed=102, n_boxes=[4, 0, 9]), 1),
    (dict(add_context_box=True, detach_appearance_embed=True), dict(B=2, img_h=128, img_w=190, seed=103, n_boxes=[0, 6]), 0),
    (dict(add_context_box=True), dict(B=2, img_h=800, img_w=1333, seed=104), 1),
])
def test_chain_is_bit_identical_to_per_kernel_orchestration(cfg_kw, batch_kw, flag, monkeypatch):
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(**batch_kw)
    monkeypatch.setattr(engine, "CHAIN", True)
    assert engine.chain_applicable("stuGuided")
    a = run_engine(cfg_kw, sd, bi, im, feats, flag)
    assert getattr(a["model"].teacher._last, "chain", False), "the native chain did not run"
    monkeypatch.setattr(engine, "CHAIN", False)
    b = run_engine(cfg_kw, sd, bi, im, feats, flag)
    assert not getattr(b["model"].teacher._last, "chain", False)
    assert a["loss"] == b["loss"]
    for k in a["tea"]:
        assert torch.equal(a["tea"][k], b["tea"][k]), k
        ga, gb = a["gfeat"][k], b["gfeat"][k]
        assert (ga is None) == (gb is None), k
        if ga is not None:
            assert torch.equal(ga, gb), k
    for l in range(len(a["masks"])):
        assert torch.equal(torch.cat(a["masks"][l], 0), torch.cat(b["masks"][l], 0))
    for n, ga in a["gparam"].items():
        gb = b["gparam"][n]
        assert (ga is None) == (gb is None), n
        if ga is not None:
            assert torch.equal(ga, gb), n


def test_chain_profile_records_every_call():
    """per-call device timing inside the native chains (bench.py's roofline source)"""
    from lgd_b200 import _lib
    from tests.gpu_util import make_model
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(2, 120, 150, seed=7)
    m = make_model(dict(add_context_box=True), sd, 1)
    f = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    _lib.profile = []
    try:
        tea, _, _, loss = m.forward(bi, im, f)
        torch.autograd.backward([loss] + list(tea.values()), [torch.ones_like(loss)] + [torch.ones_like(t) * 1e-3 for t in tea.values()])
        recs = engine.drain_profile()
    finally:
        _lib.profile = None
    names = [n for n, _ in recs]
    # eight 3x3 convolutions, forward / dgrad / wgrad each -- seven with tap rendering, where local_inst_proj_2D is
    # evaluated from per-box tap vectors (one lgd_tap_render_fwd / _bwd call instead of its three convolution launches)
    n_conv = 8 - int(engine.TAP_RENDER)
    assert names.count("lgd_tap_render_fwd") == names.count("lgd_tap_render_bwd") == int(engine.TAP_RENDER)
    assert names.count("lgd_conv3x3_fwd_f16") == n_conv and (names.count("lgd_conv3x3_dgrad_f16") +
                                                               names.count("lgd_conv3x3_dgrad_f16_gnsums") +
                                                               names.count("lgd_conv3x3_dgrad_f16_gnsums_y")) == n_conv
    assert names.count("lgd_conv3x3_wgrad_f16") == n_conv
    assert all(ms >= 0 for _, ms in recs)


def test_channels_last_boundary_is_bit_identical_and_stays_channels_last():
    """FPN maps and cotangents in channels_last memory format take the streaming gather / cast
    (no transposition) and the feature gradients come back as channels_last views of the dgrad output -- the values
    are exactly those of the NCHW boundary."""
    from tests.gpu_util import make_model
    sd = synth.synth_state_dict(5)
    bi, im, feats = synth.synth_batch(2, 200, 264, seed=21)
    outs = {}
    for layout in ("nchw", "channels_last"):
        m = make_model(dict(add_context_box=True), sd, 1)
        f = {}
        for k, v in feats.items():
            v = v.cuda()
            if layout == "channels_last":
                v = v.contiguous(memory_format=torch.channels_last)
            f[k] = v.requires_grad_(True)
        tea, _, _, loss = m.forward(bi, im, f)
        cot = synth.synth_cotangents({k: v.detach().cpu() for k, v in tea.items()})
        cot = {k: (v.cuda().contiguous(memory_format=torch.channels_last) if layout == "channels_last" else v.cuda())
               for k, v in cot.items()}
        keys = list(tea.keys())
        torch.autograd.backward([loss] + [tea[k] for k in keys], [torch.ones_like(loss)] + [cot[k] for k in keys])
        torch.cuda.synchronize()
        outs[layout] = (float(loss), {k: tea[k].detach().contiguous().cpu() for k in keys},
                        {k: f[k].grad for k in keys}, {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None})
    a, b = outs["nchw"], outs["channels_last"]
    assert a[0] == b[0]
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k])
        assert torch.equal(a[2][k].cpu(), b[2][k].contiguous().cpu())
        if b[2][k].shape[-1] * b[2][k].shape[-2] > 1:
            assert b[2][k].is_contiguous(memory_format=torch.channels_last), "feature gradients stay channels_last"
    for n in a[3]:
        assert torch.equal(a[3][n], b[3][n]), n
