"""Similarity kernel alone (tcgen05 modes) at the CUB shape for a list of batch sizes: CUDA-graph timing, optionally
through ANOTHER build of the library (A/B against an older .so) -- measurement tooling, not product code.

    python scripts/sim_only.py [--lib path.so] [--batches 64,256,1024] [--modes fp32,bf16] [--once]
"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protopformer_b200 import ops, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--batches", default="64,256,1024")
ap.add_argument("--modes", default="fp32,bf16")
ap.add_argument("--once", action="store_true", help="launch each case once and exit (for ncu)")
ap.add_argument("--K", type=int, default=81)
ap.add_argument("--P", type=int, default=2000)
ap.add_argument("--D", type=int, default=192)
a = ap.parse_args()
dev = torch.device("cuda:0")
K, P, D = a.K, a.P, a.D
new = _lib.load()
other = None
if a.lib:
    other = C.CDLL(a.lib)
    other.pph_similarity_fwd.argtypes = _lib.SIGNATURES["pph_similarity_fwd"]
    other.pph_similarity_fwd.restype = C.c_int
    other.pph_last_error_string.restype = C.c_char_p


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


for B in [int(x) for x in a.batches.split(",")]:
    g = torch.Generator(device="cpu").manual_seed(0)
    tokens = torch.randn(B, 197, D, generator=g).to(dev)
    scores = torch.rand(B, 196, generator=g).to(dev)
    Pm, Pgm = torch.rand(P, D, generator=g).to(dev), torch.rand(P, D, generator=g).to(dev)
    Wa = (torch.randn(D, D, generator=g) * (2.0 / D) ** 0.5).to(dev); ba = torch.zeros(D, device=dev)
    flops = B * (2.0 * K * D * P + 2.0 * D * P)
    with torch.no_grad():
        idx = ops.select_topk(scores, K)
        tf = ops.addon(tokens, idx, Wa, ba, True)
        pl, pg = ops.prepare_prototypes(Pm, True), ops.prepare_prototypes(Pgm, True)
        row = dict(B=B, lib=os.path.basename(a.lib) if a.lib else "current")
        for mode in a.modes.split(","):
            cfg = ops.HeadConfig(K=K, mode=mode)
            if a.once:                      # exactly one launch per case (ncu capture)
                ops._similarity_raw(cfg, tf, pl, pg)
                torch.cuda.synchronize()
                continue
            ref = ops._similarity_raw(cfg, tf, pl, pg)
            if other is not None:
                _lib._lib = other
            try:
                out = ops._similarity_raw(cfg, tf, pl, pg)
                us = graph_time(lambda: ops._similarity_raw(cfg, tf, pl, pg))
            finally:
                _lib._lib = new
            row[f"{mode}_us"] = round(us, 2)
            row[f"{mode}_tflops"] = round(flops / us / 1e6, 1)
            row[f"{mode}_same"] = bool(torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]))
        print(json.dumps(row), flush=True)
