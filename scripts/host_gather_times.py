"""Device time of the two ways to bring one host batch to the GPU: full cudaMemcpyAsync vs selection-first gather."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from protopformer_b200 import ops, synth  # noqa: E402
from protopformer_b200.graph import GraphedHeadStep  # noqa: E402

dev = torch.device("cuda:0")
s = synth.SHAPES["cub_b64"]
case = synth.make_case(s, seed=1)
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode="fp32", ppc_cov_thresh=s.ppc_cov_thresh, ppc_mean_thresh=s.ppc_mean_thresh)
params = {k: case[k].to(dev).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
st = GraphedHeadStep(params, cfg, B=s.B, N=s.N, C=s.C, m=s.m, train=False)
host = [{k: synth.make_case(s, seed=5 + i)[k].pin_memory() for k in ("tokens", "scores", "labels")} for i in range(4)]


def timed(fn, n=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n


full = timed(lambda i: st.load(0, **host[i % 4]))
print(f"full copy: {full:.1f} us  ({4 * host[0]['tokens'].numel() / full / 1e3:.1f} GB/s)")
for nc in (16, 32, 48, 64, 128, 256):
    t = timed(lambda i: st.load_host(0, host[i % 4]["tokens"], host[i % 4]["scores"], host[i % 4]["labels"], n_ctas=nc))
    moved = 4 * s.B * (s.K + 1) * s.Din
    print(f"selected rows, {nc:4d} CTAs: {t:.1f} us  ({moved / t / 1e3:.1f} GB/s over the bus)")
