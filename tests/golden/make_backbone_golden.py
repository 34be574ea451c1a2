"""Fixture for the backbone-side loop (tools/deit_models_attn.py:205-241): the reference's OWN
``MyVisionTransformer.forward_feature_mask_train_direct`` (and through it its ``attn_rollout``) driven over seeded fake
blocks (tests/util.FakeDeit), run in the build container.    python tests/golden/make_backbone_golden.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from tests.util import FakeCait, FakeDeit  # noqa: E402

# name: (B, N patches, dim, heads, depth, reserve_layer_nums, seed)
CASES = {
    "backbone_loop_small": (3, 36, 32, 2, 6, [(4, 16)], 1),
    "backbone_loop_two_stage": (2, 49, 24, 3, 7, [(3, 25), (5, 9)], 2),
}


def main():
    ref_harness.import_reference()
    import tools.deit_models_attn as dm
    torch.set_num_threads(1)
    for name, (B, N, dim, heads, depth, reserve, seed) in CASES.items():
        net = FakeDeit(dim, heads, depth)
        net.attn_rollout = types.MethodType(dm.MyVisionTransformer.attn_rollout, net)
        g = torch.Generator().manual_seed(100 + seed)
        cls_embed, x_embed = torch.randn(B, 1, dim, generator=g), torch.randn(B, N, dim, generator=g)
        with torch.no_grad():
            x, (score, _) = dm.MyVisionTransformer.forward_feature_mask_train_direct(net, cls_embed, x_embed, None, reserve)
        keep = reserve[-1][1]
        srt = score.sort(dim=-1, descending=True)[0]
        gap = ((srt[:, keep - 1] - srt[:, keep]) / srt[:, keep - 1].clamp_min(1e-30)).min().item()
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"),
                            x=x.numpy(), score=score.numpy(), sel_gap=np.float32(gap))
        print(name, tuple(x.shape), tuple(score.shape), "selection gap %.2e" % gap)
    cait()


# name: (B, N patches, dim, heads, patch depth, token-only depth, reserve_layer_nums, seed)   -- cait_models_attn.py:310-343
CAIT_CASES = {
    "backbone_loop_cait": (3, 36, 32, 2, 5, 2, [(1, 16)], 3),
}


def cait():
    import tools.cait_models_attn as cm
    for name, (B, N, dim, heads, depth, depth_t, reserve, seed) in CAIT_CASES.items():
        net = FakeCait(dim, heads, depth, depth_t)
        net.attn_rollout_cait = types.MethodType(cm.MyCait.attn_rollout_cait, net)
        g = torch.Generator().manual_seed(100 + seed)
        cls_embed, x_embed = torch.randn(B, 1, dim, generator=g), torch.randn(B, N, dim, generator=g)
        with torch.no_grad():
            x, (score, _) = cm.MyCait.forward_feature_mask_train_direct(net, cls_embed, x_embed, None, reserve)
        keep = reserve[-1][1]
        srt = score.sort(dim=-1, descending=True)[0]
        gap = ((srt[:, keep - 1] - srt[:, keep]) / srt[:, keep - 1].clamp_min(1e-30)).min().item()
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"),
                            x=x.numpy(), score=score.numpy(), sel_gap=np.float32(gap))
        print(name, tuple(x.shape), tuple(score.shape), "selection gap %.2e" % gap)


if __name__ == "__main__":
    main()
