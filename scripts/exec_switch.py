"""Where do the ~11 us between the single-buffer replay (141 us) and bench.py's rotating-slot step (152 us) go: cold inputs, or
switching between 16 graph execs?  Times (a) one exec replayed, (b) 16 execs over ONE set of input buffers, (c) 16 execs over 16.
Measured: 142.6 / 148.9 / 150.6 us -- the rotation costs 6 us per step, cold inputs 1.7 us.  Recording all 16 slot steps as ONE
graph did not recover it (151.2 us in bench.py): the cost follows the distinct kernel nodes, not the executable graph."""
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


def build(n_slots, alias):
    params = {k: case[k].to(dev).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
    for k in ("Wa", "ba", "P", "Pg"):
        params[k].requires_grad_(True)
    st = GraphedHeadStep(params, cfg, B=s.B, N=s.N, C=s.C, m=s.m, n_slots=n_slots)
    if alias:
        for i in range(1, n_slots):
            st.tokens[i], st.scores[i], st.labels[i] = st.tokens[0], st.scores[0], st.labels[0]
    for i in range(n_slots):
        c = synth.make_case(s, seed=100 + i % 4)
        st.load(i, c["tokens"], c["scores"], c["labels"])
    torch.cuda.synchronize()
    st.capture()
    return st


def timed(st, n_slots, steps=2000):
    for i in range(50):
        st.run(i % n_slots)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        st.run(i % n_slots)
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / steps


print(f"(a) 1 exec, 1 buffer set        : {timed(build(1, False), 1):.1f} us/step")
print(f"(b) 16 execs, 1 buffer set      : {timed(build(16, True), 16):.1f} us/step")
print(f"(c) 16 execs, 16 buffer sets    : {timed(build(16, False), 16):.1f} us/step")
