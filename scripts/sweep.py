"""BASELINE.json config 5: prototype-head-only sweep (tokens 49-196, prototypes 1000-8000, dim 192/384, batch 32-1024).
Times the tcgen05 similarity kernel (both precision modes) and the fused training step with CUDA events over CUDA-graph
replays and reports images/s plus the tensor-roofline fraction of the similarity kernel.  Synthetic inputs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protopformer_b200 import synth
from protopformer_b200 import ops
from protopformer_b200.graph import GraphedHeadStep

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
PEAK = float(peaks["bf16_tflops"])


def graph_time(fn, rep=20, outer=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(outer):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (rep * outer)


rows = []
cfgs = [(64, 81, 2000, 192), (256, 81, 2000, 192), (1024, 81, 2000, 192), (32, 49, 1000, 192), (256, 121, 2000, 192),
        (256, 196, 4000, 192), (128, 81, 8000, 192), (256, 81, 1200, 384), (1024, 196, 8000, 192), (512, 144, 4000, 384)]
if len(sys.argv) > 1:      # e.g. "64,81,2000,192;1024,81,2000,192"
    cfgs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[1].split(";")]
for B, K, P, D in cfgs:
    s = synth.HeadShape(f"B{B}K{K}P{P}D{D}", B, 196, D, D, K, P, P, P // 10)
    g = torch.Generator(device="cpu").manual_seed(0)
    tokens = torch.randn(B, 197, D, generator=g).to(dev)
    scores = torch.rand(B, 196, generator=g).to(dev)
    labels = torch.randint(0, s.C, (B,), generator=g).to(dev)
    Pm, Pgm = torch.rand(P, D, generator=g).to(dev), torch.rand(P, D, generator=g).to(dev)
    Wa = (torch.randn(D, D, generator=g) * (2.0 / D) ** 0.5).to(dev); ba = torch.zeros(D, device=dev)
    Wl = torch.full((s.C, P), -0.5, device=dev); Wg = torch.full((s.C, P), -0.5, device=dev)
    row = dict(B=B, K=K, P=P, D=D)
    flops = B * (2.0 * K * D * P + 2.0 * D * P)
    with torch.no_grad():
        idx = ops.select_topk(scores, K)
        tf = ops.addon(tokens, idx, Wa, ba, True)
        pl, pg = ops.prepare_prototypes(Pm, True), ops.prepare_prototypes(Pgm, True)
        for mode in ("fp32", "bf16"):
            cfg = ops.HeadConfig(K=K, mode=mode)
            us = graph_time(lambda: ops._similarity_raw(cfg, tf, pl, pg))
            row[f"sim_{mode}_us"] = round(us, 2)
            row[f"sim_{mode}_tflops"] = round(flops / us / 1e6, 1)
            row[f"sim_{mode}_frac"] = round(flops / us / 1e6 / PEAK, 4)
    params = dict(Wa=Wa, ba=ba, P=Pm, Pg=Pgm, Wl=Wl, Wg=Wg)
    for k in ("Wa", "ba", "P", "Pg"):
        params[k].requires_grad_(True)
    for mode in ("fp32", "bf16"):
        cfg = ops.HeadConfig(K=K, mode=mode)
        step = GraphedHeadStep(params, cfg, B=B, N=196, C=s.C, m=10, n_slots=1)
        step.load(0, tokens, scores, labels)
        torch.cuda.synchronize()
        step.capture()
        for _ in range(5):
            step.run(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            step.run(0)
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 50
        row[f"step_{mode}_us"] = round(us, 1)
        row[f"step_{mode}_img_s"] = round(B / us * 1e6)
        del step
    rows.append(row)
    print(json.dumps(row), flush=True)
