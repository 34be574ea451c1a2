"""Live pin: the CPU restatements against the UNMODIFIED reference imported from /root/reference (build container only;
skipped wherever the reference tree is absent, e.g. on the GPU box).  Complements the committed fixtures with fresh
seeds, so the oracle is held to the reference beyond the inputs the fixtures froze."""
import types

import pytest
import torch

from oracle import protohead_oracle as O
from oracle import ref_harness, rollout_oracle as R
from protopformer_b200 import synth
from tests.util import FakeCait, FakeDeit, norm_rel, rel_close

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("key,seed,mode,fn", [("tiny", 11, "init", "log"), ("small", 12, "init", "linear"),
                                              ("small", 13, "matched", "log")])
def test_head_oracle_matches_live_reference(key, seed, mode, fn):
    shape = synth.SHAPES[key]
    case = synth.make_case(shape, seed=seed, proto_mode=mode)
    torch.manual_seed(0)
    ref = ref_harness.run_reference(case, shape, fn=fn)
    tol = 1e-3 if mode == "matched" else 1e-4
    out = O.head_forward(case, shape.K, shape.global_coe, fn)
    assert torch.equal(out["idx"], ref["idx"])
    for k in ("logits", "act_l", "dmin_l"):
        assert rel_close(out[k], ref[k], tol), k
    tr = O.head_train_step(case, shape, fn=fn, route=ref["argmax"])
    for k in ("ce", "ppc_cov", "ppc_mean", "loss"):
        assert rel_close(tr[k], ref[k], tol), k
    for k in ("g_P", "g_Pg", "g_Wa", "g_tokens"):
        assert norm_rel(tr[k].reshape(ref[k].shape), ref[k]) < 5 * tol, k


@pytest.mark.parametrize("L,B,H,T,seed,fusion", [(2, 2, 2, 9, 21, "mean"), (5, 2, 3, 40, 22, "min"), (3, 1, 4, 64, 23, "max")])
def test_rollout_oracle_matches_live_reference(L, B, H, T, seed, fusion):
    ref_harness.import_reference()
    import tools.deit_models_attn as dm
    attn = R.synth_attention(L, B, H, T, seed)
    if not R.threshold_tie_free(attn, 0.9, fusion):
        pytest.skip("tie at the discard threshold: ATen's choice is unspecified")
    full = dm.MyVisionTransformer.attn_rollout(None, [a.clone() for a in attn], discard_ratio=0.9, head_fusion=fusion)
    assert rel_close(R.rollout_full(attn, 0.9, fusion), full, 1e-6, 1e-9)
    assert rel_close(R.rollout_cls_row(attn, 0.9, fusion), full[:, 0, 1:], 1e-5, 1e-9)


def test_cait_rollout_oracle_matches_live_reference():
    ref_harness.import_reference()
    import tools.cait_models_attn as cm
    attn = R.synth_cait_attention(4, 2, 2, 3, 20, seed=31)
    assert R.threshold_tie_free(attn, 0.9, "mean")
    _, cls_result = cm.MyCait.attn_rollout_cait(None, [a.clone() for a in attn], discard_ratio=0.9, head_fusion="mean",
                                                layer_nums=[4, 2])
    assert rel_close(R.rollout_cait(attn, 4), cls_result[:, 0], 1e-6, 1e-9)


def test_backbone_loops_match_live_reference_methods():
    """protopformer_b200/backbone.py (host logic, oracle rollout injected) against the reference's own methods."""
    from protopformer_b200.backbone import forward_feature_mask_train_direct, forward_feature_mask_train_direct_cait
    ref_harness.import_reference()
    import tools.cait_models_attn as cm
    import tools.deit_models_attn as dm

    def rollout(all_attn, topk=0, want_int64=False):
        s = R.rollout_cls_row(all_attn)
        idx = torch.topk(s, k=topk, dim=-1)[1].sort(dim=-1)[0]
        return s, idx.int(), idx

    def select(scores, K, want_int64=False):
        idx = torch.topk(scores, k=K, dim=-1)[1].sort(dim=-1)[0]
        return idx.int(), idx

    g = torch.Generator().manual_seed(77)
    net = FakeDeit(24, 2, 5)
    net.attn_rollout = types.MethodType(dm.MyVisionTransformer.attn_rollout, net)
    cls_embed, x_embed = torch.randn(2, 1, 24, generator=g), torch.randn(2, 25, 24, generator=g)
    with torch.no_grad():
        xr, (sr, _) = dm.MyVisionTransformer.forward_feature_mask_train_direct(net, cls_embed, x_embed, None, [(3, 9)])
        xo, (so, _) = forward_feature_mask_train_direct(net, cls_embed, x_embed, None, [(3, 9)], rollout=rollout)
    assert rel_close(so, sr, 1e-5, 1e-9) and rel_close(xo, xr, 1e-5, 1e-6)

    cnet = FakeCait(24, 2, 4, 2)
    cnet.attn_rollout_cait = types.MethodType(cm.MyCait.attn_rollout_cait, cnet)
    with torch.no_grad():
        xr, (sr, _) = cm.MyCait.forward_feature_mask_train_direct(cnet, cls_embed, x_embed, None, [(1, 9)])
        xo, (so, _) = forward_feature_mask_train_direct_cait(cnet, cls_embed, x_embed, None, [(1, 9)],
                                                             rollout_cait=lambda a, pre: R.rollout_cait(a, pre), select=select)
    assert rel_close(so, sr, 1e-5, 1e-9) and rel_close(xo, xr, 1e-5, 1e-6)
