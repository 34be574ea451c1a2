"""Run a few eager (no CUDA graph) training steps of the head -- the target of ncu captures.

    python scripts/run_step.py [shape=cub_b64] [mode=fp32] [impl=v2] [steps=3] [B override]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from protopformer_b200 import ops, synth  # noqa: E402
from protopformer_b200.dist import FlatGradReducer  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "cub_b64"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
impl = sys.argv[3] if len(sys.argv) > 3 else "v2"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = synth.SHAPES[key]
if len(sys.argv) > 5:
    s = s.with_batch(int(sys.argv[5]))
dev = torch.device("cuda:0")
case = {k: v.to(dev) for k, v in synth.make_case(s, seed=1).items()}
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode=mode, ppc_cov_thresh=s.ppc_cov_thresh,
                     ppc_mean_thresh=s.ppc_mean_thresh)
cls = ops.FusedHeadStep if impl == "v2" else ops.FusedHeadStepV1
kw = {}                    # the shipped default schedule (late PPC branch, gather backward, tcgen05 add-on kernels)
f = cls(cfg, s.B, s.N, s.Din, s.D, s.P, s.Pg, s.C, s.m, dev, **kw)
params = {k: case[k] for k in ("Wa", "ba", "P", "Pg")}
red = FlatGradReducer([(k, params[k]) for k in ("P", "Pg", "Wa", "ba")])
grads = dict(zip(("P", "Pg", "Wa", "ba"), red.views))
with torch.no_grad():
    for _ in range(steps):
        f.step(case["tokens"], case["scores"], case["labels"], case["Wa"], case["ba"], case["P"], case["Pg"], case["Wl"],
               case["Wg"], grads)
torch.cuda.synchronize()
print("losses", f.losses.tolist())
