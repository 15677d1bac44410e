"""Development tool: why is the pipelined end-to-end loop slower than compute + nothing? Measures the pinned-host ->
device copy of one step's FPN maps alone and underneath a running step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lgd_b200 import synth  # noqa: E402
from lgd_b200.dist import FlatGradBucket  # noqa: E402
from lgd_b200.step import HotPathDistillator  # noqa: E402

dev = torch.device("cuda", 0)
model = HotPathDistillator(synth.make_cfg(device="cuda", add_context_box=True))
model.load_hot_path_state_dict(synth.synth_state_dict(0))
model = model.to(dev)
bucket = FlatGradBucket(model.parameters())
bi, im, feats = synth.synth_batch(16, 800, 1333, seed=1234)
host = {k: v.pin_memory() for k, v in feats.items()}
res = {k: v.to(dev) for k, v in feats.items()}
buf = {k: torch.empty_like(v, device=dev) for k, v in feats.items()}
cot = {k: v.to(dev) for k, v in synth.synth_cotangents({k: torch.empty_like(v) for k, v in feats.items()}).items()}
nbytes = sum(v.numel() * 4 for v in feats.values())
cs = torch.cuda.Stream(dev)


def step():
    f = {k: v.detach().requires_grad_(True) for k, v in res.items()}
    bucket.zero_()
    model.step(bi, im, f, cot)


def copy_on(stream):
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k, v in host.items():
            buf[k].copy_(v, non_blocking=True)
        b.record()
    return a, b


for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(2):
    a, b = copy_on(cs)
    torch.cuda.synchronize()
    print("copy alone: %.2f ms = %.1f GB/s" % (a.elapsed_time(b), nbytes / a.elapsed_time(b) / 1e6))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print("step alone: %.2f ms" % (e0.elapsed_time(e1) / 5))
e0.record()
evs = []
for _ in range(5):
    evs.append(copy_on(cs))
    step()
e1.record()
torch.cuda.synchronize()
print("step with concurrent copy: %.2f ms/step; copies: %s ms" % (e0.elapsed_time(e1) / 5,
                                                                  ["%.1f" % a.elapsed_time(b) for a, b in evs]))
os.system("nvidia-smi topo -m | head -8; numactl -H 2>/dev/null | head -5; nproc")

# ---- variants of the pipelined loop
stage_buf = [{k: torch.empty_like(v, device=dev) for k, v in feats.items()} for _ in range(3)]
loss_host = [torch.empty(1).pin_memory() for _ in range(2)]


def run(label, wait_free=True, read_loss=True, copy_first=True, steps=12, nbuf=2):
    free = [None] * nbuf
    tl = []

    def stage(i):
        j = i % nbuf
        with torch.cuda.stream(cs):
            if wait_free and free[j] is not None:
                cs.wait_event(free[j])
            c0 = torch.cuda.Event(enable_timing=True)
            c0.record(cs)
            for k, v in host.items():
                stage_buf[j][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(cs)
            tl.append(("copy%d" % i, c0, ev))
        return stage_buf[j], ev, j

    def do(i, st):
        fb, ev, j = st
        torch.cuda.current_stream().wait_event(ev)
        s0 = torch.cuda.Event(enable_timing=True)
        s0.record()
        f = {k: v.detach().requires_grad_(True) for k, v in fb.items()}
        bucket.zero_()
        _, loss = model.step(bi, im, f, cot)
        d = torch.cuda.Event(enable_timing=True)
        d.record()
        free[j] = d
        tl.append(("step%d" % i, s0, d))
        return loss

    torch.cuda.synchronize()
    cs.wait_stream(torch.cuda.current_stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = stage(0)
    pending = None
    for i in range(steps):
        if copy_first:
            nxt = stage(i + 1)
            loss = do(i, st)
        else:
            loss = do(i, st)
            nxt = stage(i + 1)
        st = nxt
        if read_loss:
            slot = loss_host[i % 2]
            slot.copy_(loss.detach().reshape(1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[1].synchronize()
                float(pending[0][0])
            pending = (slot, ev)
    e1.record()
    torch.cuda.synchronize()
    print("%-50s %.2f ms/step" % (label, e0.elapsed_time(e1) / steps))
    print("   timeline:", "  ".join("%s[%.1f-%.1f]" % (n, e0.elapsed_time(a), e0.elapsed_time(b)) for n, a, b in tl[:14]))


run("pipelined: wait_free, read_loss, copy_first")
run("pipelined: 3 buffers, wait_free, read_loss", nbuf=3)
run("pipelined: 3 buffers, wait_free, no read_loss", nbuf=3, read_loss=False)
