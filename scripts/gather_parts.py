"""Device time of the sparse backward's pieces alone: parts = 1 (token rows + CLS rows), 2 (prototype rows), 3 (both)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from protopformer_b200 import _lib, ops, synth  # noqa: E402
from protopformer_b200.graph import GraphedHeadStep  # noqa: E402

dev = torch.device("cuda:0")
s = synth.SHAPES["cub_b64"]
case = synth.make_case(s, seed=1)
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode="fp32", ppc_cov_thresh=s.ppc_cov_thresh, ppc_mean_thresh=s.ppc_mean_thresh)
params = {k: case[k].to(dev).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
for k in ("Wa", "ba", "P", "Pg"):
    params[k].requires_grad_(True)
st = GraphedHeadStep(params, cfg, B=s.B, N=s.N, C=s.C, m=s.m)
st.load(0, case["tokens"], case["scores"], case["labels"])
st.capture()
st.run(0)
torch.cuda.synchronize()
f = st.fused
gP, gPg = torch.empty_like(params["P"]), torch.empty_like(params["Pg"])


def timed(fn, per=16, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / (per * reps)


for parts in (1, 2, 3):
    t = timed(lambda: _lib.call("pph_similarity_bwd_fused", parts, f.g_l, f.g_g, f.argmin, f.Zs, f.Zc, params["P"].detach(),
                                params["Pg"].detach(), s.B, s.K, s.D, s.P, s.Pg, s.m, f.ws_gather, f.ws_bins, None, None, 1,
                                f.dZs, f.dZc, gP, gPg))
    print(f"parts={parts}: {t:.1f} us")

# token rows alone (no global prototypes -> no CLS slices): which piece is the long pole of parts = 1?
t = timed(lambda: _lib.call("pph_similarity_bwd_fused", 1, f.g_l, None, f.argmin, f.Zs, None, params["P"].detach(), None, s.B, s.K,
                            s.D, s.P, 0, s.m, f.ws_gather, f.ws_bins, None, None, 1, f.dZs, None, gP, None))
print(f"token rows only: {t:.1f} us")
