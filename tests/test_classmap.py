"""Class-row activation maps on the original token grid (eval_interpretability.py:195-225).

The consumer is script code (argparse + dataset loading at import time), so it cannot be imported: the oracle restates
its gather (:198-202) and scatter (:218-223) lines and is checked here against an index loop (CPU) and fed the
REFERENCE's own `proto_acts` from the golden fixtures (GPU test)."""
import numpy as np
import pytest
import torch

from oracle import protohead_oracle as O
from tests.util import load_golden, rel_close, max_rel


def test_oracle_gather_scatter_matches_index_loop():
    g = torch.Generator().manual_seed(0)
    B, P, K, m, N = 3, 12, 4, 3, 9
    act = torch.rand(B, P, K, generator=g)
    idx = torch.stack([torch.randperm(N, generator=g)[:K].sort()[0] for _ in range(B)])
    labels = torch.tensor([0, 3, 2])
    got = O.class_activation_maps(act, idx, labels, m, N)
    want = torch.zeros(B, m, N)
    for b in range(B):
        for q in range(m):
            for k in range(K):
                want[b, q, idx[b, k]] = act[b, int(labels[b]) * m + q, k]
    assert torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_s1", "tiny_s2_linear", "small_s1", "small_s3_matched"])
def test_class_maps_match_reference_proto_acts(name):
    from protopformer_b200 import ops
    shape, case, g, fn = load_golden(name)
    dev = torch.device("cuda:0")
    d = {k: v.to(dev) for k, v in case.items()}
    cfg = ops.HeadConfig(K=shape.K, global_coe=shape.global_coe, act_fn=fn, mode="fp32_fma")
    idx = ops.select_topk(d["scores"], shape.K)
    tf = ops.addon(d["tokens"], idx, d["Wa"], d["ba"], False)
    maps = ops.class_activation_maps(cfg, tf, d["P"], d["labels"], shape.m, shape.N)
    side = int(round(shape.N ** 0.5))
    assert maps.shape == (shape.B, shape.m, side, side)
    # the reference's own push_forward activations (fixture), gathered / scattered as eval_interpretability.py does
    want = O.class_activation_maps(torch.from_numpy(g["act_map"]), torch.from_numpy(g["idx"]).long(), case["labels"],
                                   shape.m, shape.N)
    tol = 1e-3 if "matched" in name else 1e-4
    assert rel_close(maps.flatten(2).cpu(), want, tol), max_rel(maps.flatten(2).cpu(), want)
    assert int((maps.flatten(2) != 0).sum(-1).max()) <= shape.K          # pruned tokens stay exactly zero


@pytest.mark.gpu
def test_class_maps_full_size_against_materialised_map():
    from protopformer_b200 import synth
    from protopformer_b200 import ops
    shape = synth.SHAPES["cub_b64"].with_batch(16)
    case = synth.make_case(shape, seed=2)
    dev = torch.device("cuda:0")
    d = {k: v.to(dev) for k, v in case.items()}
    cfg = ops.HeadConfig(K=shape.K, mode="fp32_fma")
    idx = ops.select_topk(d["scores"], shape.K)
    tf = ops.addon(d["tokens"], idx, d["Wa"], d["ba"], False)
    maps = ops.class_activation_maps(cfg, tf, d["P"], d["labels"], shape.m, shape.N)
    _, act_map = ops.materialize_maps(cfg, tf, d["P"], d["Pg"])                    # (B,P,K) from the FP32-FMA kernel
    want = O.class_activation_maps(act_map.cpu(), idx.cpu().long(), case["labels"], shape.m, shape.N)
    assert rel_close(maps.flatten(2).cpu(), want, 1e-4), max_rel(maps.flatten(2).cpu(), want)


@pytest.mark.gpu
def test_class_maps_variant_2_in_subprocess():
    import os
    import subprocess
    import sys
    env = dict(os.environ, PPH_CLASSMAP="2")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-m", "gpu", "-x", "-k", "not subprocess"], env=env, capture_output=True,
                       text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:]
