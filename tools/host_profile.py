"""Development tool: where does the HOST time of one distillation step go? (cProfile over a few steps, B=16.)"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lgd_b200 import synth  # noqa: E402
from lgd_b200.dist import FlatGradBucket  # noqa: E402
from lgd_b200.step import HotPathDistillator  # noqa: E402

B = int(os.environ.get("B", "16"))
dev = torch.device("cuda", 0)
model = HotPathDistillator(synth.make_cfg(device="cuda", add_context_box=True))
model.load_hot_path_state_dict(synth.synth_state_dict(0))
model = model.to(dev)
bucket = FlatGradBucket(model.parameters())
bi, im, feats = synth.synth_batch(B, 800, 1333, seed=1234)
res = {k: v.to(dev) for k, v in feats.items()}
cot = {k: v.to(dev) for k, v in synth.synth_cotangents({k: torch.empty_like(v) for k, v in feats.items()}).items()}


def step():
    f = {k: v.detach().requires_grad_(True) for k, v in res.items()}
    bucket.zero_()
    model.step(bi, im, f, cot)


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step, total %.2f ms/step" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(25)
