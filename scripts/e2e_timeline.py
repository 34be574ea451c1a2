"""CUPTI timeline of the end-to-end loop of bench.py (selection-first host transfer on a copy stream, step graph on
the main stream, loss read back every step): which kernels overlap and where the step waits.

    python scripts/e2e_timeline.py [n_ctas=48] [slots=2]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from protopformer_b200 import ops, synth  # noqa: E402
from protopformer_b200.graph import GraphedHeadStep  # noqa: E402

n_ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nslot = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
s = synth.SHAPES["cub_b64"]
case = synth.make_case(s, seed=1)
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode="fp32", ppc_cov_thresh=s.ppc_cov_thresh, ppc_mean_thresh=s.ppc_mean_thresh)
params = {k: case[k].to(dev).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
for k in ("Wa", "ba", "P", "Pg"):
    params[k].requires_grad_(True)
nbuf = 16
st = GraphedHeadStep(params, cfg, B=s.B, N=s.N, C=s.C, m=s.m, n_slots=nslot)
host = []
for i in range(nbuf):
    c = synth.make_case(s, seed=100 + i % 4)
    host.append({k: c[k].pin_memory() for k in ("tokens", "scores", "labels")})
for i in range(nslot):
    st.load(i, host[i]["tokens"], host[i]["scores"], host[i]["labels"])
torch.cuda.synchronize()
st.capture()
EXP = os.environ.get("EXP", "")
from protopformer_b200 import _lib  # noqa: E402


def custom_load(sl, h):
    """Variants of load_host for the diagnosis: 'nocopy' = no memcpy nodes, 'norows' = no gather kernel."""
    with torch.no_grad():
        if "nocopy" not in EXP:
            st.scores[sl].copy_(h["scores"], non_blocking=True)
            st.labels[sl].copy_(h["labels"], non_blocking=True)
        if "nosel" not in EXP:
            _lib.call("pph_select_topk", st.scores[sl], s.B, 1, s.N, s.K, st._load_idx[sl], None)
        if "norows" not in EXP:
            _lib.call("pph_gather_rows_host", h["tokens"].data_ptr(), st._load_idx[sl], s.B, s.N, s.Din, s.K,
                      st.tokens[sl].detach(), n_ctas)


def cap(sl, h):
    st.load_host(sl, h["tokens"], h["scores"], h["labels"], n_ctas)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        custom_load(sl, h)
    return g


lg = {(sl, i): cap(sl, host[i]) for sl in range(nslot) for i in range(nbuf) if i % nslot == sl}
copy_stream = torch.cuda.Stream()
out_host = torch.zeros(1).pin_memory()
ready = [torch.cuda.Event() for _ in range(nslot)]
done = [torch.cuda.Event() for _ in range(nslot)]
for e in done:
    e.record(torch.cuda.current_stream())


def loop(n):
    cur = torch.cuda.current_stream()
    for i in range(n):
        sl = i % nslot
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[sl])
            lg[(sl, i % nbuf)].replay()
            ready[sl].record(copy_stream)
        cur.wait_event(ready[sl])
        st.run(sl)
        if "nod2h" not in EXP:
            out_host.copy_(st.loss[sl].reshape(1), non_blocking=True)
        done[sl].record(cur)
    torch.cuda.synchronize()


loop(64)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
loop(1600)
b.record()
torch.cuda.synchronize()
print(f"# EXP={EXP!r} n_ctas {n_ctas}, slots {nslot}: {1e3 * a.elapsed_time(b) / 1600:.1f} us per end-to-end step")
if os.environ.get("PPH_TIMELINE"):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        loop(8)
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    sel = [i for i, e in enumerate(ev) if "select_topk" in e.name]
    lo = sel[6] if len(sel) > 6 else 0                      # two selections per step (transfer + step): start of step 3
    t0 = ev[lo].time_range.start
    for e in ev[lo:lo + 40]:
        print(f"{e.time_range.start - t0:9.2f} us  +{e.time_range.end - e.time_range.start:8.2f} us  {e.name[:60]}")
