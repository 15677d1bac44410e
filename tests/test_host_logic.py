"""CPU: host-side logic of the plugin surface -- the box table (a1, host half), registries, state_dict contract,
error behaviour without a GPU -- against the oracle and the reference's naming contract."""
import pytest
import torch

import lgd_b200
from lgd_b200 import engine, synth
from lgd_b200.step import HotPathDistillator
from oracle import lgd_oracle as O


@pytest.mark.parametrize("ctx", [True, False])
def test_box_table_matches_oracle(ctx):
    bi, im, _ = synth.synth_batch(4, 120, 150, seed=7, adversarial=True, n_boxes=[None, 0, 5, 64])
    H, W = im.tensor.shape[-2:]
    tb = engine.build_box_table(bi, H, W, ctx, "cpu")
    ref = O.prepare_boxes([x["instances"] for x in bi], H, W, ctx)
    assert tb.counts == [b.shape[0] for b, _, _ in ref]
    assert torch.equal(tb.boxes.view(-1, 4), torch.cat([b for b, _, _ in ref], 0))       # clamped boxes, bit exact
    onehot = torch.cat([oh for _, oh, _ in ref], 0)
    lab = tb.labels.long()
    assert torch.equal(onehot.sum(1) > 0, lab >= 0)                                        # ctx / dummy rows: no class
    assert torch.equal(onehot.argmax(1)[lab >= 0], lab[lab >= 0])
    assert tb.img_start.tolist()[-1] == tb.T == sum(tb.counts)
    for i, (il, (_, _, ril)) in enumerate(zip(tb.inst_labels, ref)):
        assert torch.equal(il.float(), ril.float()), i
    # image 1 has no GT: single dummy row, never a context box appended (label_encoder.py:57-69,75)
    assert tb.counts[1] == 1 and tb.n_render.tolist()[1] == (0 if ctx else 1)
    if ctx:
        assert tb.ctx_row.tolist() == [s + n - 1 for s, n in zip(tb.img_start.tolist()[:-1], tb.counts)]
    else:
        assert tb.ctx_row.tolist() == [-1] * 4


def test_labels_out_of_range_assert():
    bi, im, _ = synth.synth_batch(1, 64, 64, seed=1, n_boxes=[2])
    bi[0]["instances"].gt_classes = torch.tensor([3, 80])
    with pytest.raises(AssertionError):
        engine.build_box_table(bi, 64, 64, True, "cpu")


def test_registries_resolve_reference_names():
    for n in ("DistillatorRetinaNet", "DistillatorGeneralizedRCNN", "DistillatorFCOS", "DistillatorPOTO",
              "DistillatorATSS"):
        assert issubclass(lgd_b200.META_ARCH_REGISTRY.get(n), lgd_b200.BaseDistillator)
    assert lgd_b200.CUSTOMIZED_DETECTORS_REGISTRY.get("DynamicTeacher") is lgd_b200.DynamicTeacher
    assert lgd_b200.ADAPTERS_REGISTRY.get("SequentialConvs") is lgd_b200.SequentialConvs


def test_state_dict_contract():
    m = HotPathDistillator(synth.make_cfg())
    own = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = synth.hot_path_param_shapes()
    assert own == want                      # names AND shapes of teacher.* / adapter.distill.* (no buffers)
    assert sum(v.numel() for k, v in m.state_dict().items() if k.startswith("teacher.")) == 8304336
    assert sum(v.numel() for k, v in m.state_dict().items() if k.startswith("adapter.")) == 1770240
    m.load_hot_path_state_dict(synth.synth_state_dict(3))


@pytest.mark.parametrize("cfg_kw", [dict(category_format="norm_classes"), dict(load_labelmap=True),
                                    dict(category_format="norm_classes", load_labelmap=True)])
def test_state_dict_contract_of_the_other_descriptor_formats(cfg_kw):
    """label_encoder.py:136-145: the descriptor length (5 with norm_classes, +49 with LOAD_LABELMAP) sets the shapes of
    stn_desc (k x k transform) and conv1; everything else is unchanged. Checked against the unmodified reference's own
    state_dict by the golden cases noctx_stu_normcls / seg_ctx_detach (their weights load into both)."""
    d = synth.desc_dim_of(cfg_kw)
    assert d == {(True, False): 5, (False, True): 133, (True, True): 54}[
        (cfg_kw.get("category_format") == "norm_classes", bool(cfg_kw.get("load_labelmap")))]
    m = HotPathDistillator(synth.make_cfg(**cfg_kw))
    own = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert own == synth.hot_path_param_shapes(d)
    assert own["teacher.label_encoder_.stn_desc.fc3.weight"] == (d * d, 256)
    assert own["teacher.label_encoder_.conv1.weight"] == (64, d, 1)
    m.load_hot_path_state_dict(synth.synth_state_dict(3, desc_dim=d))
    with pytest.raises(ValueError):
        HotPathDistillator(synth.make_cfg(category_format="bogus"))


def test_no_cpu_fallback_and_error_behaviour():
    m = HotPathDistillator(synth.make_cfg())
    bi, im, feats = synth.synth_batch(1, 64, 64, seed=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.teacher((bi, im, None, feats))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.distill_loss({"stu": feats, "tea": feats}, im, bi, None, None)
    bad = HotPathDistillator(synth.make_cfg(interact_pattern="bogus"))
    with pytest.raises(ValueError):
        bad.teacher((bi, im, None, feats))


def test_pyramid_geometry_and_synth_workload():
    assert synth.pyramid_hw(800, 1344) == [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    assert sum(h * w for h, w in synth.pyramid_hw(800, 1344)) == 22400
    assert synth.pyramid_hw(640, 1088)[-1] == (5, 9)


def test_seg_level_masks_match_the_reference_restatement():
    """LOAD_LABELMAP host half (engine.seg_level_masks) == get_segmask_inside_gt as restated by the oracle
    (rasterise at image size, background row, zero padding, F.interpolate nearest) -- bit exact, with and without
    the background row, including an image without GT and images smaller than the padded batch."""
    import torch
    from lgd_b200 import engine, synth
    from oracle import lgd_oracle as O
    for ih, iw, unp in [(128, 160, [(128, 160), (120, 150), (100, 160)]), (400, 666, [(400, 666), (390, 600), (320, 550)])]:
        bi, im, feats = synth.synth_batch(3, ih, iw, seed=105, n_boxes=[4, 0, 6], with_masks=True, unpadded=unp)
        Hp, Wp = im.tensor.shape[-2:]
        hws = [tuple(v.shape[-2:]) for v in feats.values()]
        for ctx in (True, False):
            got = engine.seg_level_masks(bi, Hp, Wp, hws, ctx)
            ref = O.seg_inside_masks(hws, bi, (Hp, Wp), ctx, synth.polygons_to_bitmask)
            flat = torch.cat([torch.cat(lv, 0).reshape(-1) for lv in ref])
            assert torch.equal(got.float(), flat)
    tb = engine.build_box_table(bi, Hp, Wp, True, "cpu", with_mask_descriptors=True)
    assert tuple(tb.mask49.shape) == (tb.T, 49)
    assert bool((tb.mask49[4] == 1).all()) and bool((tb.mask49[5] == 0).all())     # context row ones, dummy row zeros
