"""Kernel timeline of one CUDA-graph replay of the training step (CUPTI activity records through torch.profiler):
start offset, duration and stream of every kernel -- shows which branches really overlap.

    python scripts/timeline.py [shape=cub_b64] [mode=fp32] [impl=v2] [key=value variants ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from protopformer_b200 import ops, synth  # noqa: E402
from protopformer_b200.graph import GraphedHeadStep  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "cub_b64"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
impl = sys.argv[3] if len(sys.argv) > 3 else "v2"
variants = dict(a.split("=") for a in sys.argv[4:]) or None
dev = torch.device("cuda:0")
s = synth.SHAPES[key]
case = synth.make_case(s, seed=1)
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode=mode, ppc_cov_thresh=s.ppc_cov_thresh,
                     ppc_mean_thresh=s.ppc_mean_thresh)
params = {k: case[k].to(dev).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
for k in ("Wa", "ba", "P", "Pg"):
    params[k].requires_grad_(True)
st = GraphedHeadStep(params, cfg, B=s.B, N=s.N, C=s.C, m=s.m, train=True, impl=impl, variants=variants)
st.load(0, case["tokens"], case["scores"], case["labels"])
torch.cuda.synchronize()
st.capture()
for _ in range(20):
    st.run(0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        st.run(0)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
per = len(ev) // 6
one = ev[4 * per:5 * per]              # the fifth replay
t0 = one[0].time_range.start
print(f"# {impl} {variants} : {per} device activities per replay")
for e in one:
    print(f"{e.time_range.start - t0:9.2f} us  +{e.time_range.end - e.time_range.start:8.2f} us  {e.name[:70]}")
print(f"# replay span {one[-1].time_range.end - t0:.2f} us; next replay starts at {ev[5 * per].time_range.start - t0:.2f} us")
