"""Live pin: the restated oracle against the reference's own source executed verbatim (only where the read-only
checkout is mounted, i.e. the build container; skipped on the GPU box)."""
import pytest
import torch

from oracle import reference_fn as R
from oracle import verbatim
from wseg_b200 import synth

pytestmark = pytest.mark.skipif(not verbatim.available(), reason="reference checkout not mounted")


@pytest.mark.parametrize("canonical", [True, False])
def test_loop_body_equals_restatement(canonical):
    n, hw1, hw2 = 2, (160, 176), (128, 128)
    v1 = synth.head_outputs(n, hw1, seed=41)
    v2 = synth.head_outputs(n, hw2, seed=42)
    lab = synth.labels(n, 43)

    def leaves(d):
        return {k: v.clone().requires_grad_(True) for k, v in d.items()}

    a1, a2, b1, b2 = leaves(v1), leaves(v2), leaves(v1), leaves(v2)
    stub = verbatim.HeadStub()
    stub.add(hw1, a1["cam"], a1["cam_rv"], a1["f_proj"], a1["cam_rv_down"])
    stub.add(hw2, a2["cam"], a2["cam_rv"], a2["f_proj"], a2["cam_rv_down"])
    ns = verbatim.run_loop_body(stub, torch.zeros(n, 3, *hw1), lab, 0.2, py_seed=3, canonical_ties=canonical)
    ns["loss"].backward()
    out = R.contrast_step_loss(b1["cam"], b1["cam_rv"], b1["f_proj"], b1["cam_rv_down"], b2["cam"], b2["cam_rv"],
                               b2["f_proj"], b2["cam_rv_down"], lab, 0.2, hw2, sampler=R.ReferenceSampler(3),
                               topk_fn=R.canonical_topk if canonical else torch.topk)
    out["loss"].backward()
    for k in R.LOSS_KEYS:
        assert float(ns[k]) == pytest.approx(float(out[k]), rel=1e-6, abs=1e-9), k
    assert torch.equal(ns["pseudo_label1"], out["pseudo_label1"])
    assert torch.equal(ns["pseudo_label2"], out["pseudo_label2"])
    assert torch.equal(ns["top_indices"], out["top_indices1"])
    assert torch.equal(ns["top_indices2"], out["top_indices2"])
    for k in ("cam", "cam_rv", "f_proj"):
        for a, b in ((a1, b1), (a2, b2)):
            assert torch.allclose(a[k].grad, b[k].grad, rtol=1e-5, atol=1e-9), k


def test_net_forward_equals_restatement():
    mod = verbatim.import_reference_net()
    torch.manual_seed(5)
    net = mod.Net()
    net.eval()                       # the reference train() returns None, so no chaining
    cap = {}
    orig = net.PCM
    net.PCM = lambda cam, f: cap.setdefault("out", orig(cap.setdefault("cam", cam), cap.setdefault("f", f)))
    hook = net.fc8.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o))
    with torch.no_grad():
        net(torch.randn(1, 3, 64, 64))
    hook.remove()
    assert torch.equal(R.cam_d_norm(cap["logits"]), cap["cam"])
    with torch.no_grad():
        assert torch.allclose(R.pcm(cap["cam"], net.f9(cap["f"])), cap["out"], atol=2e-6)
