"""Multi-tensor optimizer steps (lgd_b200/optim.py -> lgd_mt_sgd / lgd_mt_adamw) against torch.optim on the same
device, with the reference's grouping: one parameter group per parameter (utils/build.py:497-508)."""
import pytest
import torch

from lgd_b200 import synth
from lgd_b200.optim import FusedAdamW, FusedSGD, build_distillator_optimizer, reduce_loss_dict

pytestmark = pytest.mark.gpu

SHAPES = [(7056, 256), (256, 256, 3, 3), (64,), (1, ), (1024, 128, 1), (33000,)]


def _params(seed):
    gen = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=gen).cuda()) for s in SHAPES]


def _groups(ps, lr, wd):
    return [{"params": [p], "lr": lr, "weight_decay": wd} for p in ps]


def _run(opt_a, opt_b, pa, pb, steps, skip_last_grad_at=None):
    gen = torch.Generator().manual_seed(99)
    for it in range(steps):
        for i, (a, b) in enumerate(zip(pa, pb)):
            g = torch.randn(a.shape, generator=gen).cuda() * 0.1
            if skip_last_grad_at == it and i == len(pa) - 1:
                a.grad = b.grad = None      # a parameter without gradient this step is skipped (torch semantics)
            else:
                a.grad, b.grad = g.clone(), g.clone()
        if it == 2:                          # an LR scheduler changes the groups' lr between steps
            for grp in opt_a.param_groups + opt_b.param_groups:
                grp["lr"] *= 0.5
        opt_a.step()
        opt_b.step()
    torch.cuda.synchronize()


def test_fused_sgd_matches_torch():
    pa, pb = _params(1), _params(1)
    a = FusedSGD(_groups(pa, 0.01, 1e-4), 0.01, momentum=0.9)
    b = torch.optim.SGD(_groups(pb, 0.01, 1e-4), 0.01, momentum=0.9)
    _run(a, b, pa, pb, 5, skip_last_grad_at=1)
    for x, y in zip(pa, pb):
        assert float((x - y).norm() / y.norm()) < 1e-6
        assert float((a.state[x]["momentum_buffer"] - b.state[y]["momentum_buffer"]).norm()) <= 1e-6 * float(b.state[y]["momentum_buffer"].norm())
    # state_dict layout is torch's (checkpoint compatible)
    assert set(a.state_dict()["state"][0].keys()) == set(b.state_dict()["state"][0].keys())


def test_fused_adamw_matches_torch():
    pa, pb = _params(2), _params(2)
    a = FusedAdamW(_groups(pa, 1e-3, 0.05), 1e-3)
    b = torch.optim.AdamW(_groups(pb, 1e-3, 0.05), 1e-3, betas=(0.9, 0.999))
    _run(a, b, pa, pb, 6)
    for x, y in zip(pa, pb):
        assert float((x - y).norm() / y.norm()) < 2e-6


def test_build_distillator_optimizer_and_loss_reduce():
    from types import SimpleNamespace as ns
    from lgd_b200.step import HotPathDistillator
    cfg = synth.make_cfg(device="cuda")
    solver = ns(OPTIMIZER="SGD", BASE_LR=0.01, MOMENTUM=0.9, WEIGHT_DECAY=1e-4)
    cfg.MODEL.DISTILLATOR.STUDENT.SOLVER = solver
    cfg.MODEL.DISTILLATOR.TEACHER.SOLVER = ns(OPTIMIZER="ADAMW", BASE_LR=1e-4, MOMENTUM=0.9, WEIGHT_DECAY=0.05)
    model = HotPathDistillator(cfg).cuda()
    stu_opt, tea_opt = build_distillator_optimizer(cfg, ns(module=model))
    assert isinstance(stu_opt, FusedSGD) and isinstance(tea_opt, FusedAdamW)
    n_tea = sum(1 for _ in model.teacher.parameters())
    n_stu = sum(1 for _ in model.student.parameters()) + sum(1 for _ in model.adapter.parameters())
    assert len(tea_opt.param_groups) == n_tea and len(stu_opt.param_groups) == n_stu      # one group per parameter
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    stu_opt.step()
    tea_opt.step()
    for n, p in model.named_parameters():
        assert not torch.equal(p, before[n]), n
    out = reduce_loss_dict({"loss_b": torch.tensor(2.0).cuda(), "loss_a": torch.tensor(1.5).cuda()})
    assert out == {"loss_a": 1.5, "loss_b": 2.0}


def test_publish_scalars_writes_pinned_host_memory():
    """lgd_store_to_host: the step's losses land in pinned host memory by a kernel store (no copy engine); pageable
    memory is refused."""
    from lgd_b200.optim import publish_scalars
    vals = torch.tensor([1.5, -2.25, 3.0e-7], device="cuda")
    slot = torch.zeros(4).pin_memory()
    ev = publish_scalars(vals, slot)
    ev.synchronize()
    assert slot[:3].tolist() == vals.cpu().tolist() and slot[3] == 0
    with pytest.raises(ValueError):
        publish_scalars(vals, torch.zeros(4))
