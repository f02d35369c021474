"""The restated oracle (oracle/reference_fn.py) against golden vectors produced by the reference's own code
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_fn as R
from oracle.make_golden import LOSS_CASE, NET_CASE, PCM_CASE, checksum, sample
from wseg_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_synthetic_inputs_reproduce(golden_dir):
    g = _load(golden_dir, "loss_n2_192_128.npz")
    c = LOSS_CASE
    v1 = synth.head_outputs(c["n"], c["view1"], seed=c["seed1"])
    v2 = synth.head_outputs(c["n"], c["view2"], seed=c["seed2"])
    got1 = np.stack([checksum(v1[k]) for k in ("cam", "cam_rv", "f_proj", "cam_rv_down")])
    got2 = np.stack([checksum(v2[k]) for k in ("cam", "cam_rv", "f_proj", "cam_rv_down")])
    np.testing.assert_allclose(got1, g["in_sum_v1"], rtol=1e-12)
    np.testing.assert_allclose(got2, g["in_sum_v2"], rtol=1e-12)
    np.testing.assert_array_equal(synth.labels(c["n"], c["label_seed"]).numpy(), g["label"])


def test_loss_body_matches_reference(golden_dir):
    g = _load(golden_dir, "loss_n2_192_128.npz")
    c = LOSS_CASE
    v1 = {k: v.requires_grad_(True) for k, v in synth.head_outputs(c["n"], c["view1"], seed=c["seed1"]).items()}
    v2 = {k: v.requires_grad_(True) for k, v in synth.head_outputs(c["n"], c["view2"], seed=c["seed2"]).items()}
    lab = synth.labels(c["n"], c["label_seed"])
    out = R.contrast_step_loss(v1["cam"], v1["cam_rv"], v1["f_proj"], v1["cam_rv_down"], v2["cam"], v2["cam_rv"],
                               v2["f_proj"], v2["cam_rv_down"], lab, c["bg_threshold"], c["view2"],
                               sampler=R.ReferenceSampler(c["py_seed"]))
    out["loss"].backward()
    for k in R.LOSS_KEYS:
        assert float(out[k]) == pytest.approx(float(g[k]), rel=1e-6), k
    for v in "12":
        np.testing.assert_array_equal(out["pseudo_label" + v].numpy(), g["pseudo_label" + v])
        np.testing.assert_array_equal(out["top_indices" + v].numpy(), g["top_indices" + v])
        np.testing.assert_allclose(out["prototypes" + v].numpy(), g["prototypes" + v], atol=1e-7)
        np.testing.assert_allclose(out["norm_cam" + v].numpy(), g["norm_cam" + v], atol=1e-7)
    np.testing.assert_allclose(v1["f_proj"].grad.numpy(), g["grad_f_proj1"], atol=1e-9, rtol=1e-5)
    np.testing.assert_allclose(v2["f_proj"].grad.numpy(), g["grad_f_proj2"], atol=1e-9, rtol=1e-5)
    np.testing.assert_allclose(sample(v1["cam"].grad), g["grad_cam1_sample"], atol=1e-9, rtol=1e-5)
    np.testing.assert_allclose(sample(v1["cam_rv"].grad), g["grad_cam_rv1_sample"], atol=1e-9, rtol=1e-5)
    np.testing.assert_allclose(sample(v2["cam_rv"].grad), g["grad_cam_rv2_sample"], atol=1e-9, rtol=1e-5)
    assert v1["cam_rv_down"].grad is None and v2["cam_rv_down"].grad is None     # seeds carry no gradient


def test_pcm_matches_reference(golden_dir):
    g = _load(golden_dir, "pcm_n2_20x13.npz")
    c = PCM_CASE
    d = synth.pcm_inputs(c["n"], c["h"], c["w"], c["seed"])
    camn = R.cam_d_norm(d["cam"])
    np.testing.assert_array_equal(camn.numpy(), g["cam_d_norm"])
    f = d["f"].clone().requires_grad_(True)
    out = R.pcm(camn, f)
    out.backward(d["grad_out"])
    np.testing.assert_allclose(out.detach().numpy(), g["cam_rv"], atol=1e-7)
    np.testing.assert_allclose(f.grad.numpy(), g["grad_f"], atol=1e-10, rtol=1e-4)


def test_network_head_matches_reference(golden_dir):
    g = _load(golden_dir, "net_head_72x56.npz")
    camn = R.cam_d_norm(torch.from_numpy(g["cam_logits"]))
    np.testing.assert_array_equal(camn.numpy(), g["cam_d_norm"])
    f9 = torch.nn.functional.conv2d(torch.from_numpy(g["f"]), torch.from_numpy(g["f9_weight"]))
    out = R.pcm(camn, f9)
    np.testing.assert_allclose(out.numpy(), g["cam_rv_down"], atol=2e-6)
    assert NET_CASE["size"] == (72, 56)


def test_canonical_topk_is_stable():
    x = torch.tensor([[0.2, 0.5, 0.2, 0.5, 0.1, 0.2]])
    v, i = R.canonical_topk(x, 4)
    assert i.tolist() == [[1, 3, 0, 2]]
    assert torch.equal(v, torch.tensor([[0.5, 0.5, 0.2, 0.2]]))
    v, i = R.canonical_topk(x[0], 3, largest=False)
    assert i.tolist() == [4, 0, 2]
