"""Helpers shared by the parity tests: load tests/golden/*.npz and rebuild their inputs."""
import os

import numpy as np
import torch

from lgd_b200 import synth
from oracle.make_golden import CASES

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    cfg_kw, batch_kw, flag = CASES[name]
    sd = synth.synth_state_dict(int(g["weight_seed"]), desc_dim=synth.desc_dim_of(cfg_kw))
    wsum = np.array([float(v.double().sum()) for _, v in sorted(sd.items())])
    assert np.allclose(wsum, g["wsum"], rtol=0, atol=1e-9), "synthetic weights drifted from golden"
    bi, im, feats = synth.synth_batch(**batch_kw)
    for k, v in feats.items():
        assert abs(float(v.double().sum()) - float(g[f"feat_sum_{k}"])) < 1e-6, "synthetic features drifted"
    return g, cfg_kw, batch_kw, flag, sd, bi, im, feats


def unpack_mask(g, key):
    shape = g[f"mask_shape_{key}"]
    return torch.from_numpy(np.unpackbits(g[f"mask_{key}"], axis=1)[:, :shape[1]].astype(np.float32))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_head_case(name):
    """tests/golden/<name>.npz of oracle.make_golden.HEAD_CASES: the reference's FCOSHead / POTOHead outputs and gradients
    on seeded inputs. Returns (golden dict, case tuple, state dict, features, cotangents)."""
    from oracle.make_golden import HEAD_CASES, head_inputs
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    sd, feats, cots = head_inputs(name)
    wsum = np.array([float(v.double().sum()) for _, v in sorted(sd.items())])
    assert np.allclose(wsum, g["wsum"], rtol=0, atol=1e-9), "synthetic head weights drifted from golden"
    assert np.allclose(np.array([float(f.double().sum()) for f in feats]), g["feat_sum"], rtol=0, atol=1e-6)
    return g, HEAD_CASES[name], sd, feats, cots
