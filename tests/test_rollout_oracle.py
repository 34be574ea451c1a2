"""CPU suite: the rollout restatement (oracle/rollout_oracle.py) against the fixtures produced by the reference's own
``attn_rollout`` (tests/golden/make_rollout_golden.py), and the algebra the CUDA path relies on."""
import os

import numpy as np
import pytest
import torch

from oracle import rollout_oracle as R
from protopformer_b200 import synth
from tests.util import GOLDEN_DIR, rel_close

ROLLOUT_CASES = {
    "rollout_tiny": (3, 2, 2, 12, 1, "mean", 3),
    "rollout_small_h6": (4, 3, 6, 50, 2, "mean", 25),
    "rollout_max_fusion": (3, 2, 4, 30, 3, "max", 9),
    "rollout_deit_tiny_b2": (11, 2, 3, 197, 4, "mean", 81),
}


def load_rollout(name):
    L, B, H, T, seed, fusion, K = ROLLOUT_CASES[name]
    attn = R.synth_attention(L, B, H, T, seed)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    chk = np.array([synth.checksum(a) for a in attn])
    assert np.allclose(chk, g["chk"], rtol=1e-9), "synthetic attention drifted from the fixture"
    return attn, g, fusion, K


@pytest.mark.parametrize("name", list(ROLLOUT_CASES))
def test_full_product_restatement_is_the_reference(name):
    attn, g, fusion, K = load_rollout(name)
    full = R.rollout_full(attn, 0.9, fusion)
    assert rel_close(full[:, 0, 1:], g["scores"], 1e-6, 1e-9)
    assert rel_close(full[:, 0].sum(-1), g["row_sum"], 1e-6)


@pytest.mark.parametrize("name", list(ROLLOUT_CASES))
def test_cls_row_chain_equals_row_zero_of_the_product(name):
    attn, g, fusion, K = load_rollout(name)
    row = R.rollout_cls_row(attn, 0.9, fusion)
    assert rel_close(row, g["scores"], 1e-5, 1e-9)
    idx = torch.topk(row, k=K, dim=-1)[1].sort(dim=-1)[0]
    assert np.array_equal(idx.numpy(), g["idx"])          # fixture's selection gap >> the chain's rounding noise
    assert float(g["sel_gap"]) > 1e-4


def test_discard_count_and_tie_rule():
    assert R.discard_count(197 * 197, 0.9) == 34928       # int(38809 * 0.9) as in deit_models_attn.py:110
    x = torch.tensor([[[3.0, 1.0], [1.0, 2.0]]])
    kept = R.discard_smallest(x, 1)                       # two entries equal the threshold: lowest flat index goes
    assert kept.flatten().tolist() == [3.0, 0.0, 1.0, 2.0]
    assert R.discard_smallest(x, 0).equal(x) and R.discard_smallest(x, 4).abs().sum() == 0
    assert not R.threshold_tie_free([x.unsqueeze(1)], 0.25)


def test_start_row_variant_matches_manual_chain():
    """CaiT-style start row (cait_models_attn.py:255-259): v0 @ a_{L-1} ... a_0 with all columns kept."""
    attn = R.synth_attention(3, 2, 2, 10, seed=7)
    v0 = torch.rand(2, 10, generator=torch.Generator().manual_seed(0))
    got = R.rollout_cls_row(attn, v0=v0, drop_first=False)
    full = R.rollout_full(attn)
    assert rel_close(got, torch.einsum("bi,bij->bj", v0, full), 1e-5, 1e-9)


CAIT_CASES = {
    "rollout_cait_tiny": (3, 1, 2, 2, 16, 1, 2),
    "rollout_cait_xxs24_b2": (24, 2, 2, 4, 196, 2, 121),
}


def load_cait(name):
    n_patch, n_cls, B, H, T, seed, K = CAIT_CASES[name]
    attn = R.synth_cait_attention(n_patch, n_cls, B, H, T, seed)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    chk = np.array([synth.checksum(a) for a in attn])
    assert np.allclose(chk, g["chk"], rtol=1e-9), "synthetic attention drifted from the fixture"
    return attn, g, n_patch, K


@pytest.mark.parametrize("name", list(CAIT_CASES))
def test_cait_restatement_and_start_row_chain_match_the_reference(name):
    """tools/cait_models_attn.py:223-261 through the reference's own attn_rollout_cait (fixture)."""
    attn, g, n_patch, K = load_cait(name)
    assert rel_close(R.rollout_cait(attn, n_patch), g["scores"], 1e-6, 1e-9)
    v0 = torch.cat([R.process_layer(a) for a in attn[n_patch:]], dim=1).mean(dim=1)[:, 1:]
    chain = R.rollout_cls_row(attn[:n_patch], v0=v0, drop_first=False)
    assert rel_close(chain, g["scores"], 1e-5, 1e-9)
    assert np.array_equal(torch.topk(chain, k=K, dim=-1)[1].sort(dim=-1)[0].numpy(), g["idx"])
