"""Per-launch and whole-step timings of the round-2 step next to the round-1 sequence (CUDA graphs, CUDA events).

    python scripts/step_times.py [shape=cub_b64] [mode=fp32] [B override]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from protopformer_b200 import _lib, ops, synth  # noqa: E402
from protopformer_b200.graph import GraphedHeadStep  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "cub_b64"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
s = synth.SHAPES[key]
if len(sys.argv) > 3:
    s = s.with_batch(int(sys.argv[3]))
case = synth.make_case(s, seed=1)
cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode=mode, ppc_cov_thresh=s.ppc_cov_thresh,
                     ppc_mean_thresh=s.ppc_mean_thresh)
REP = 20


def timed_graph(fn, rep=REP, outer=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(outer):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (outer * rep)


def make(impl, train=True, variants=None, stop_after=0):
    params = {k: case[k].to(dev).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
    for k in ("Wa", "ba", "P", "Pg"):
        params[k].requires_grad_(train)
    st = GraphedHeadStep(params, cfg, B=s.B, N=s.N, C=s.C, m=s.m, train=train, impl=impl, variants=variants)
    st.load(0, case["tokens"], case["scores"], case["labels"])
    if stop_after:
        st.fused.stop_after = stop_after
    torch.cuda.synchronize()
    st.capture()
    return st, params


out = {"shape": key, "B": s.B, "mode": mode}
S = dict(prep="simt", addon_bwd="simt", bwd="staged", ppc="inline")
T = dict(prep="tc", addon_bwd="tc", bwd="gather", ppc="late")
ARMS = [("v1", None), ("v2", S), ("v2", T), ("v2", dict(T, ppc="inline")), ("v2", dict(T, ppc="split")), ("v2", dict(T, bwd="staged", ppc="split")),
        ("v2", dict(T, prep="simt")), ("v2", dict(T, addon_bwd="simt"))]
for impl, variants in ARMS:
    for train in (True, False):
        if not train and variants is not None and variants not in (S, T):
            continue
        tag = impl if variants is None else "v2[" + ",".join(f"{k}={v}" for k, v in variants.items()) + "]"
        try:
            st, _ = make(impl, train, variants)
        except Exception as exc:
            out[f"{tag}_{'train' if train else 'eval'}"] = f"failed: {exc}"
            continue
        for _ in range(20):
            st.run(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300):
            st.run(0)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 300
        out[f"{tag}_{'train' if train else 'eval'}"] = {"us": round(us, 2), "launches": st.kernel_launches_per_step}
        if impl == "v2" and train and variants == S:
            f = st.fused
            p = st.p
            grads = st.grads
            tok, sc, lab = st.tokens[0].detach(), st.scores[0], st.labels[0]
            B, N, Din, D, Pn, Pgn, C, m, H = f.dims
            K = cfg.K
            c = _lib.call
            per = {}
            with torch.no_grad():
                per["head_prep"] = timed_graph(lambda: c(
                    "pph_head_prep", sc, tok, p["Wa"], p["ba"], B, max(H, 1), N, Din, D, K, 0.5, f.idx32, None, f.Zs, f.Zc,
                    f.z2s, f.z2c, f.z2s_ctr, f.z2c_ctr, f.z2s_hi, f.z2c_hi, f.Zs_hi, f.Zs_lo, f.Zc_hi, f.Zc_lo,
                    p["P"], Pn, f.P_hi, f.P_lo, f.p2, f.p2_ctr, f.p2_hi, p["Pg"], Pgn, f.Pg_hi, f.Pg_lo, f.pg2, f.pg2_ctr,
                    f.pg2_hi))
                sel = {_lib.MODE_FP32_FMA: 0, _lib.MODE_BF16X3: 1, _lib.MODE_BF16: 2}[cfg.mode_id]
                per["similarity_fwd"] = timed_graph(lambda: c(
                    "pph_similarity_fwd", cfg.mode_id, cfg.act_id, float(cfg.eps), B, K, D, Pn, Pgn, f.Zs, f.Zc,
                    (f.z2s, f.z2s_ctr, f.z2s_hi)[sel], (f.z2c, f.z2c_ctr, f.z2c_hi)[sel], f.Zs_hi, f.Zs_lo, f.Zc_hi, f.Zc_lo,
                    p["P"], p["Pg"], (f.p2, f.p2_ctr, f.p2_hi)[sel], (f.pg2, f.pg2_ctr, f.pg2_hi)[sel], f.P_hi, f.P_lo,
                    f.Pg_hi, f.Pg_lo, f.dmin_l, f.argmin, f.act_l, f.dmin_g, f.act_g, None, None))

                def mid(train_=1, ppc_=1):
                    c("pph_head_mid", f.act_l, f.act_g, f.dmin_l, f.dmin_g, f.argmin, p["Wl"], p["Wg"], lab, B, K, D, Pn, Pgn,
                      C, m, N, float(cfg.global_coe), cfg.act_id, float(cfg.eps), 1.0, train_, ppc_, f.Zs, f.z2s, p["P"],
                      f.p2, f.idx32, float(cfg.ppc_cov_thresh), float(cfg.ppc_mean_thresh), 0.1, 0.5, f.ws_mid, f.ws_bins,
                      f.logits, f.logits_g, f.logits_l, f.losses, f.dlogits, f.g_l, f.g_g, f.pairT,
                      f.dZs_ppc if ppc_ else None, f.dP_img if ppc_ else None, None)
                per["head_mid"] = timed_graph(lambda: mid(1, 1))
                per["head_mid(no ppc)"] = timed_graph(lambda: mid(1, 0))
                per["head_mid(eval)"] = timed_graph(lambda: mid(0, 0))
                per["head_mid(LL+bins+cls only, mode 2)"] = timed_graph(lambda: mid(1, 2))
                per["head_mid(PPC only, mode 3)"] = timed_graph(lambda: mid(1, 3))
                sd, e_a, e_b = torch.cuda.Stream(), torch.cuda.Event(), torch.cuda.Event()

                def mid_split():
                    cur = torch.cuda.current_stream()
                    e_a.record(cur)
                    sd.wait_event(e_a)
                    with torch.cuda.stream(sd):
                        mid(1, 3)
                        e_b.record(sd)
                    mid(1, 2)
                    cur.wait_event(e_b)
                per["head_mid(mode 2 || mode 3)"] = timed_graph(mid_split)
                def bwd2(parts):
                    c("pph_similarity_bwd2", parts, f.g_l, f.g_g, f.pairT, f.ws_bins, f.Zs, f.Zc, p["P"], p["Pg"], B, K, D, Pn,
                      Pgn, m, f.dZs_ppc, f.dP_img, 1, f.dZs, f.dZc, grads["P"], grads["Pg"])
                per["similarity_bwd2[all, serial]"] = timed_graph(lambda: bwd2(7))
                for mask, nm in ((1, "tokens"), (2, "protos"), (4, "cls")):
                    per[f"similarity_bwd2[{nm}]"] = timed_graph(lambda: bwd2(mask))

                def ab2(parts):
                    c("pph_addon_bwd2", parts, tok, f.idx32, p["Wa"], f.dZs, f.dZc, B, N, Din, D, K, f.ws_addon,
                      grads["Wa"], grads["ba"], f.dtokens)
                per["addon_bwd2"] = timed_graph(lambda: ab2(3))
                per["addon_bwd2[wgrad]"] = timed_graph(lambda: ab2(1))
                per["addon_bwd2[dgrad]"] = timed_graph(lambda: ab2(2))

                def ab3(parts):
                    c("pph_addon_bwd3", parts, tok, f.idx32, p["Wa"], f.dZs, f.dZc, None, B, N, Din, D, K, f.ws_tc,
                      grads["Wa"], grads["ba"], f.dtokens)
                per["addon_bwd3[wgrad]"] = timed_graph(lambda: ab3(1))
                per["addon_bwd3[dgrad]"] = timed_graph(lambda: ab3(2))
                per["select_topk"] = timed_graph(lambda: c("pph_select_topk", sc, B, max(H, 1), N, K, f.idx32, None))
                per["addon_fwd2"] = timed_graph(lambda: c(
                    "pph_addon_fwd2", tok, f.idx32, p["Wa"], p["ba"], B, N, Din, D, K, f.Zs, f.Zc, f.z2s, f.z2c, 0.5,
                    f.z2s_ctr, f.z2c_ctr, f.z2s_hi, f.z2c_hi, f.Zs_hi, f.Zs_lo, f.Zc_hi, f.Zc_lo, f.ws_tc))
                per = {k: round(v, 2) for k, v in per.items()}
            out["v2_per_launch_us"] = per
            out["v2_sum_of_launches_us"] = (per["head_prep"] + per["similarity_fwd"] + per["head_mid"] +
                                            per["similarity_bwd2[all, serial]"] + per["addon_bwd2"])
def time_step(st, n=300):
    for _ in range(20):
        st.run(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        st.run(0)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


cum = {}
for k, name in ((1, "prep"), (2, "+similarity"), (3, "+mid"), (4, "+bwd"), (0, "+addon_bwd = step")):
    st, _ = make("v2", True, T, stop_after=k)
    cum[name] = round(time_step(st), 2)
out["v2[T] cumulative graph time by stage"] = cum
print(json.dumps(out))
