"""Warm-cache per-entry-point timings at a BASELINE shape (CUDA graph of REP back-to-back calls, CUDA events)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from protopformer_b200 import synth
from protopformer_b200 import ops, _lib

key = sys.argv[1] if len(sys.argv) > 1 else "cub_b64"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
s = synth.SHAPES[key]
case = {k: v.to(dev) for k, v in synth.make_case(s, seed=1).items()}
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode=mode)
REP = 20


def timeit(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (10 * REP)
    print(f"{name:28s} {us:9.2f} us")
    return us

res = {}
with torch.no_grad():
    tok, sc = case["tokens"], case["scores"]
    res["select_topk"] = timeit("select_topk", lambda: ops.select_topk(sc, s.K))
    idx = ops.select_topk(sc, s.K)
    res["addon_fwd"] = timeit("addon_fwd", lambda: ops.addon(tok, idx, case["Wa"], case["ba"], True))
    tf = ops.addon(tok, idx, case["Wa"], case["ba"], True)
    res["split_rows(P)"] = timeit("split_rows(P)", lambda: ops.prepare_prototypes(case["P"], True))
    pl, pg = ops.prepare_prototypes(case["P"], True), ops.prepare_prototypes(case["Pg"], True)
    for m in ("fp32", "bf16", "fp32_fma"):
        c = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode=m)
        res[f"similarity_fwd[{m}]"] = timeit(f"similarity_fwd[{m}]", lambda: ops._similarity_raw(c, tf, pl, pg))
    dmin_l, argmin, act_l, dmin_g, act_g, _, _ = ops._similarity_raw(cfg, tf, pl, pg)
    B, P, Pg, C, K, D, N, Din = s.B, s.P, s.Pg, s.C, s.K, s.D, s.N, s.Din
    logits = torch.empty(B, C, device=dev); lg = torch.empty_like(logits); ll = torch.empty_like(logits)
    res["logits_fwd"] = timeit("logits_fwd", lambda: _lib.call("pph_logits_fwd", act_l, act_g, case["Wl"], case["Wg"], B, P, Pg, C, 0.5, logits, lg, ll))
    labels = case["labels"]
    dsl = torch.empty(B, s.m, K, device=dev); st = torch.empty(B, s.m, 8, device=dev); part = torch.empty(B, 2, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev); losses = torch.empty(2, device=dev)
    res["ppc_fwd"] = timeit("ppc_fwd", lambda: _lib.call("pph_ppc_fwd", tf.Zs, tf.z2s, pl.P, pl.p2, idx, labels, B, K, D, P, s.m, N, 0, 1e-4, 1.0, 2.0, dsl, st, part, cnt, losses))
    gl = torch.ones(2, device=dev)
    dZp = torch.empty_like(tf.Zs); dPp = torch.zeros_like(pl.P)
    res["ppc_bwd"] = timeit("ppc_bwd", lambda: _lib.call("pph_ppc_bwd", tf.Zs, pl.P, idx, labels, dsl, st, gl, 1.0, 1.0, B, K, D, P, s.m, N, 0, 1e-4, 1.0, 2.0, 1, dZp, dPp))
    dlog = torch.randn(B, C, device=dev) / B
    g_l = torch.empty(B, P, device=dev); g_g = torch.empty(B, Pg, device=dev)
    res["logits_bwd"] = timeit("logits_bwd", lambda: _lib.call("pph_logits_bwd", dlog, None, None, case["Wl"], case["Wg"], dmin_l, dmin_g, B, P, Pg, C, 0.5, 0, 1e-4, g_l, g_g))
    dZs = torch.empty_like(tf.Zs); dZc = torch.empty_like(tf.Zc); dPl = torch.empty_like(pl.P); dPg = torch.empty_like(pg.P)
    ws = ops.bwd_workspace(B, K, D, P, Pg, dev)
    res["similarity_bwd"] = timeit("similarity_bwd", lambda: _lib.call("pph_similarity_bwd", g_l, g_g, argmin, tf.Zs, tf.Zc, pl.P, pg.P, B, K, D, P, Pg, ws, 3, None, None, dZs, dZc, dPl, dPg))
    lt_part = torch.empty(B, device=dev); lt_cnt = torch.zeros(1, dtype=torch.int32, device=dev); lt_out = torch.empty(4, device=dev)
    res["loss_tail"] = timeit("loss_tail", lambda: _lib.call("pph_loss_tail", logits, labels, losses, 0.1, 0.5, 1.0, B, C, lt_part, lt_cnt, lt_out, dlog))
    dWa = torch.empty_like(case["Wa"]); dba = torch.empty(D, device=dev); dtok = torch.empty_like(tok)
    wsa = ops.addon_bwd_workspace(B, N, Din, D, K, dev)
    res["addon_bwd"] = timeit("addon_bwd", lambda: _lib.call("pph_addon_bwd", tok, idx, case["Wa"], tf.Zs, tf.Zc, dZs, dZc, B, N, Din, D, K, wsa, 3, dWa, dba, dtok))
    res["torch_cross_entropy_fwd"] = timeit("torch CE fwd", lambda: F.cross_entropy(logits, labels))
    res["torch_zeros_like(P)"] = timeit("torch zeros_like(P)", lambda: torch.zeros_like(pl.P))
    res["torch_add(Zs)"] = timeit("torch add (B,K,D)", lambda: dZs + dZp)
print(json.dumps({"shape": key, "mode": mode, "us": res}))
