"""Host-side logic that needs no GPU: reference-RNG mining keys, network surface, candidate exchange semantics."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import reference_fn as R
from wseg_b200 import loss as L
from wseg_b200 import ops


def _select_by_keys(pl, key, c):
    members = torch.nonzero(pl == c).reshape(-1)
    ln = members.numel()
    order = sorted(range(ln), key=lambda i: (int(key[members[i]]) & 0xFFFFFFFF, int(members[i])))
    return set(int(members[i]) for i in order[: ln // 2])


def test_reference_rng_keys_reproduce_python_draws():
    g = torch.Generator().manual_seed(3)
    pl1 = torch.randint(0, 21, (300,), generator=g)
    pl2 = torch.randint(0, 5, (300,), generator=g)
    pl2[7] = 20                                   # a class with a single pixel: skipped, no draw consumed
    k1, k2 = L.reference_rng_keys(pl1, pl2, 11)
    sampler = R.ReferenceSampler(11)
    for pl, key in ((pl1, k1), (pl2, k2)):
        sampler.burn_dead_draws(pl.numel())
        sim = torch.zeros(pl.numel())
        _, sets = R.hard_pixel_sets(sim, pl, 21, sampler)
        for c, rnd, _hard in sets:
            assert set(rnd.tolist()) == _select_by_keys(pl, key, c), c
    # the same object can be a shared random.Random (the driver's --reference-rng mode)
    rng = random.Random(11)
    k1b, _ = L.reference_rng_keys(pl1, pl2, rng)
    assert torch.equal(k1, k1b)


def test_hash_keys_deterministic():
    a = ops.hash_keys(1000, 5, "cpu")
    b = ops.hash_keys(1000, 5, "cpu")
    c = ops.hash_keys(1000, 6, "cpu")
    assert torch.equal(a, b) and not torch.equal(a, c) and a.dtype == torch.int32


def test_net_surface_matches_reference_checkpoint_format(golden_dir):
    from wseg_b200.network import resnet38_contrast
    net = resnet38_contrast.Net()
    want = json.load(open(os.path.join(golden_dir, "state_dict_shapes.json")))
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(got) == list(want) and got == want
    net.train()
    groups = net.get_parameter_groups()
    assert [len(g) for g in groups] == [35, 0, 5, 0]
    assert all(not m.training for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d))
    assert not net.conv1a.weight.requires_grad and not net.b2_2.conv_branch2a.weight.requires_grad
    assert net.f9.weight.requires_grad and net.f9.weight.shape == (192, 195, 1, 1)
    assert hasattr(net, "normalize") and hasattr(net, "PCM") and len(net.from_scratch_layers) == 5


def test_seam_net_surface_matches_reference_checkpoint_format(golden_dir):
    """network/resnet38_SEAM.py (the baseline network, same PCM): parameter names / shapes as the reference's."""
    from wseg_b200.network import resnet38_SEAM
    net = resnet38_SEAM.Net()
    want = json.load(open(os.path.join(golden_dir, "state_dict_shapes_seam.json")))
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(got) == list(want) and got == want
    assert [len(g) for g in net.get_parameter_groups()] == [43, 0, 4, 0]        # as the reference's, before .train()
    net.train()
    assert [len(g) for g in net.get_parameter_groups()] == [35, 0, 4, 0]
    assert len(net.from_scratch_layers) == 4 and not hasattr(net, "fc_proj")


def test_union_of_local_topk_contains_global_topk():
    """The exchange step relies on: global top-k (canonical order) is a subset of the ranks' local top-k's."""
    g = torch.Generator().manual_seed(0)
    c, p_local, world, k = 5, 64, 4, 8
    planes = [torch.randint(0, 6, (c, p_local), generator=g).float() / 5 for _ in range(world)]   # many ties
    full = torch.cat(planes, dim=1)
    gv, gi = R.canonical_topk(full, k)
    cand_v, cand_i = [], []
    for r, pl in enumerate(planes):
        v, i = R.canonical_topk(pl, k)
        cand_v.append(v)
        cand_i.append(i + r * p_local)
    cv, ci = torch.cat(cand_v, 1), torch.cat(cand_i, 1)
    order = np.lexsort((ci.numpy(), -cv.numpy()), axis=1)[:, :k]
    sel_i = np.take_along_axis(ci.numpy(), order, 1)
    assert np.array_equal(sel_i, gi.numpy())


def test_philox_known_answers():
    """The Random123 known-answer vectors of Philox4x32-10, for the host reference the GPU keys are checked against."""
    assert ops.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ops.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ops.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_poly_optimizer_reproduces_the_reference_quirks():
    """tool/torchutils.py:13-14 passes the weight decay positionally into SGD's momentum slot; the poly power is the
    constructor's `momentum`; the drop-in must expose the same param_groups (golden: the reference's own class)."""
    import numpy as np
    import os
    from oracle.make_golden import OPT_CASE, optim_groups, optim_inputs
    from wseg_b200.optim import PolyOptimizer
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "poly_optimizer.npz"))
    init, _ = optim_inputs()
    params = [torch.nn.Parameter(p.clone()) for p in init]
    opt = PolyOptimizer(optim_groups(params, OPT_CASE["lr"], OPT_CASE["wt_dec"]), lr=OPT_CASE["lr"],
                        weight_decay=OPT_CASE["wt_dec"], max_step=OPT_CASE["max_step"])
    assert [g["momentum"] for g in opt.param_groups] == list(gold["group_momentum"])
    assert [g["weight_decay"] for g in opt.param_groups] == list(gold["group_weight_decay"])
    assert opt.defaults["weight_decay"] == 0 and opt.defaults["momentum"] == OPT_CASE["wt_dec"]
    for p in params:                                     # CPU parameters with gradients: there is no CPU path
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError):
        opt.step()
