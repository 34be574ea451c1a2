"""The backbone-side loop that consumes the rollout score (protopformer_b200/backbone.py) against the reference's own
``forward_feature_mask_train_direct`` (fixtures from tests/golden/make_backbone_golden.py).  CPU: host logic with the
oracle's rollout injected; the CUDA ops it calls by default are held to the same oracle in tests/test_rollout_gpu.py."""
import os

import numpy as np
import pytest
import torch

from oracle import rollout_oracle as R
from tests.util import GOLDEN_DIR, FakeCait, FakeDeit, rel_close

CASES = {
    "backbone_loop_small": (3, 36, 32, 2, 6, [(4, 16)], 1),
    "backbone_loop_two_stage": (2, 49, 24, 3, 7, [(3, 25), (5, 9)], 2),
}


def _oracle_rollout(all_attn, topk=0, want_int64=False):
    scores = R.rollout_cls_row(all_attn)
    idx = torch.topk(scores, k=topk, dim=-1)[1].sort(dim=-1)[0]
    return scores, idx.int(), idx


@pytest.mark.parametrize("name", list(CASES))
def test_backbone_loop_matches_reference_method(name):
    from protopformer_b200.backbone import forward_feature_mask_train_direct, patch_deit_features
    B, N, dim, heads, depth, reserve, seed = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    net = FakeDeit(dim, heads, depth)
    gen = torch.Generator().manual_seed(100 + seed)
    cls_embed, x_embed = torch.randn(B, 1, dim, generator=gen), torch.randn(B, N, dim, generator=gen)
    with torch.no_grad():
        x, (score, none) = forward_feature_mask_train_direct(net, cls_embed, x_embed, None, reserve,
                                                             rollout=_oracle_rollout)
    assert none is None and float(g["sel_gap"]) > 1e-4
    assert rel_close(score, g["score"], 1e-5, 1e-9)
    assert rel_close(x, g["x"], 1e-5, 1e-6)           # same kept tokens -> same masked attention in the later blocks
    patched = patch_deit_features(FakeDeit(dim, heads, depth))
    assert patched.forward_feature_mask_train_direct.__func__ is forward_feature_mask_train_direct


def test_cait_backbone_loop_matches_reference_method():
    """tools/cait_models_attn.py:310-343 through the reference's own method (fixture), oracle rollout injected."""
    from protopformer_b200.backbone import forward_feature_mask_train_direct_cait, patch_cait_features
    B, N, dim, heads, depth, depth_t, reserve, seed = 3, 36, 32, 2, 5, 2, [(1, 16)], 3
    g = dict(np.load(os.path.join(GOLDEN_DIR, "backbone_loop_cait.npz")))
    net = FakeCait(dim, heads, depth, depth_t)
    gen = torch.Generator().manual_seed(100 + seed)
    cls_embed, x_embed = torch.randn(B, 1, dim, generator=gen), torch.randn(B, N, dim, generator=gen)

    def select(scores, K, want_int64=False):
        idx = torch.topk(scores, k=K, dim=-1)[1].sort(dim=-1)[0]
        return idx.int(), idx

    with torch.no_grad():
        x, (score, none) = forward_feature_mask_train_direct_cait(
            net, cls_embed, x_embed, None, reserve, rollout_cait=lambda a, pre: R.rollout_cait(a, pre), select=select)
    assert none is None and float(g["sel_gap"]) > 1e-4
    assert rel_close(score, g["score"], 1e-5, 1e-9)
    assert rel_close(x, g["x"], 1e-5, 1e-6)
    patched = patch_cait_features(FakeCait(dim, heads, depth, depth_t))
    assert patched.forward_feature_mask_train_direct.__func__ is forward_feature_mask_train_direct_cait


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_backbone_loop_with_cuda_rollout(name):
    """Same loop, same GPU-resident fake blocks, CUDA rollout vs the oracle's rollout injected (maps copied to the host):
    identical attention maps reach both, so scores agree to the rollout tolerance and the kept tokens are the same."""
    from protopformer_b200.backbone import forward_feature_mask_train_direct, patch_deit_features
    B, N, dim, heads, depth, reserve, seed = CASES[name]
    dev = torch.device("cuda:0")
    net = patch_deit_features(FakeDeit(dim, heads, depth).to(dev))
    gen = torch.Generator().manual_seed(100 + seed)
    cls_embed = torch.randn(B, 1, dim, generator=gen).to(dev)
    x_embed = torch.randn(B, N, dim, generator=gen).to(dev)

    def host_rollout(all_attn, topk=0, want_int64=False):
        s, i32, i64 = _oracle_rollout([a.cpu() for a in all_attn], topk, want_int64)
        return s.to(dev), i32.to(dev), i64.to(dev)

    with torch.no_grad():
        x, (score, _) = net.forward_feature_mask_train_direct(cls_embed, x_embed, None, reserve)
        x2, (score2, _) = forward_feature_mask_train_direct(net, cls_embed, x_embed, None, reserve, rollout=host_rollout)
    assert rel_close(score.cpu(), score2.cpu(), 1e-5, 1e-9)
    assert torch.equal(x, x2)
