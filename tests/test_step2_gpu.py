"""The round-2 training step (ops.FusedHeadStep: pph_select_addon_fwd / pph_head_prep -> pph_similarity_fwd -> pph_head_mid ->
pph_similarity_bwd_fused / pph_similarity_bwd2 -> pph_addon_bwd3 / pph_addon_bwd2) kernel by kernel against the CPU oracle, against the round-1 entry points on identical device inputs,
and end to end against the reference fixtures -- including the benchmarked shape (cub_b64) in both precision modes.

Tolerances: indices bit-exact; FP32-FMA kernels (prep, mid, bwd2, addon_bwd2) within 2e-5 relative of the oracle /
1e-5 of the largest entry against the round-1 kernels (same arithmetic, different summation order); whole step as in
tests/test_gpu_parity.py (1e-4 relative forward, gradients relative to the largest entry of each tensor)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import protohead_oracle as O
from protopformer_b200 import synth
from tests.util import GOLDEN_CASES, load_golden, max_rel, norm_rel, rel_close

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from protopformer_b200 import ops
    return ops


def _lib():
    from protopformer_b200 import _lib
    return _lib


def _d(case):
    return {k: v.to(DEV) for k, v in case.items()}


def _cfg(shape, mode="fp32", fn="log"):
    return _ops().HeadConfig(K=shape.K, global_coe=shape.global_coe, act_fn=fn, mode=mode,
                             ppc_cov_thresh=shape.ppc_cov_thresh, ppc_mean_thresh=shape.ppc_mean_thresh)


def _e(*s, dt=torch.float32):
    return torch.empty(s, dtype=dt, device=DEV)


def _prep(shape, d, heads=0, split=True):
    """Run pph_head_prep; returns a dict of its outputs."""
    B, N, Din, D, K, P, Pg = shape.B, shape.N, shape.Din, shape.D, shape.K, shape.P, shape.Pg
    bf, i32 = torch.bfloat16, torch.int32
    o = dict(idx32=_e(B, K, dt=i32), idx64=_e(B, K, dt=torch.int64), Zs=_e(B, K, D), Zc=_e(B, D), z2s=_e(B, K), z2c=_e(B),
             z2s_ctr=_e(B, K), z2c_ctr=_e(B), z2s_hi=_e(B, K), z2c_hi=_e(B), Zs_hi=_e(B * K, D, dt=bf),
             Zs_lo=_e(B * K, D, dt=bf), Zc_hi=_e(B, D, dt=bf), Zc_lo=_e(B, D, dt=bf), P_hi=_e(P, D, dt=bf),
             P_lo=_e(P, D, dt=bf), p2=_e(P), p2_ctr=_e(P), p2_hi=_e(P), Pg_hi=_e(Pg, D, dt=bf), Pg_lo=_e(Pg, D, dt=bf),
             pg2=_e(Pg), pg2_ctr=_e(Pg), pg2_hi=_e(Pg))
    scores = d["scores_h"] if heads else d["scores"]
    _lib().call("pph_head_prep", scores.contiguous(), d["tokens"], d["Wa"], d["ba"], B, max(heads, 1), N, Din, D, K, 0.5,
                o["idx32"], o["idx64"], o["Zs"], o["Zc"], o["z2s"], o["z2c"], o["z2s_ctr"], o["z2c_ctr"], o["z2s_hi"],
                o["z2c_hi"], o["Zs_hi"], o["Zs_lo"], o["Zc_hi"], o["Zc_lo"],
                d["P"], P, o["P_hi"], o["P_lo"], o["p2"], o["p2_ctr"], o["p2_hi"],
                d["Pg"], Pg, o["Pg_hi"], o["Pg_lo"], o["pg2"], o["pg2_ctr"], o["pg2_hi"])
    torch.cuda.synchronize()
    return o


PREP_CASES = [("tiny", 0, 1), ("small", 0, 1), ("cub_b8", 0, 1), ("cub_b64", 0, 2), ("cub_b8", 3, 3), ("cars_b64", 0, 4),
              ("dogs_b256", 0, 5), ("sweep_k49", 0, 1), ("sweep_k196", 0, 1), ("sweep_k144_d384", 6, 1)]


@pytest.mark.parametrize("key,heads,seed", PREP_CASES)
def test_head_prep_matches_oracle(key, heads, seed):
    shape = synth.SHAPES[key]
    if key == "dogs_b256":
        shape = shape.with_batch(9)
    case = synth.make_case(shape, seed=seed, heads=heads)
    o = _prep(shape, _d(case), heads)
    scores = case["scores_h"] if heads else case["scores"]
    idx = O.select_tokens(scores, shape.K)
    assert torch.equal(o["idx32"].cpu().long(), idx) and torch.equal(o["idx64"].cpu(), idx)
    Zs, Zc = O.addon(case["tokens"], idx, case["Wa"], case["ba"])
    assert rel_close(o["Zs"].cpu(), Zs, 2e-5), max_rel(o["Zs"].cpu(), Zs)
    assert rel_close(o["Zc"].cpu(), Zc, 2e-5)
    assert rel_close(o["z2s"].cpu(), (Zs * Zs).sum(-1), 1e-5) and rel_close(o["z2c"].cpu(), (Zc * Zc).sum(-1), 1e-5)
    assert rel_close(o["z2s_ctr"].cpu(), ((Zs - 0.5) ** 2).sum(-1), 1e-5)
    assert rel_close(o["z2c_ctr"].cpu(), ((Zc - 0.5) ** 2).sum(-1), 1e-5)
    rec = (o["Zs_hi"].float() + o["Zs_lo"].float()).cpu().reshape(Zs.shape)
    assert float((rec - (o["Zs"].cpu() - 0.5)).abs().max()) < 1e-5
    assert torch.equal(o["Zs_hi"].cpu().reshape(Zs.shape), (o["Zs"].cpu() - 0.5).to(torch.bfloat16))
    assert rel_close(o["z2s_hi"].cpu(), (o["Zs_hi"].float() ** 2).sum(-1).reshape(shape.B, shape.K).cpu(), 1e-5)
    assert rel_close(o["z2c_hi"].cpu(), (o["Zc_hi"].float() ** 2).sum(-1).cpu(), 1e-5)
    for name, pre in (("P", "p"), ("Pg", "pg")):
        V = case[name].reshape(case[name].shape[0], -1)
        assert torch.equal(o[name + "_hi"].cpu(), (V - 0.5).to(torch.bfloat16))
        lo = ((V - 0.5) - (V - 0.5).to(torch.bfloat16).float()).to(torch.bfloat16)
        assert torch.equal(o[name + "_lo"].cpu(), lo)
        assert rel_close(o[pre + "2"].cpu(), (V * V).sum(-1), 1e-5)
        assert rel_close(o[pre + "2_ctr"].cpu(), ((V - 0.5) ** 2).sum(-1), 1e-5)
        assert rel_close(o[pre + "2_hi"].cpu(), ((V - 0.5).to(torch.bfloat16).float() ** 2).sum(-1), 1e-5)


def test_head_prep_selection_edges_nan_and_ties():
    """NaN scores must not break the exactly-K contract (torch.topk ranks NaN first); all-equal
    scores take the lowest indices; the stand-alone pph_select_topk follows the same rule."""
    shape = synth.SHAPES["cub_b8"]
    case = synth.make_case(shape, seed=7)
    s = case["scores"].clone()
    s[0, 5] = float("nan")
    s[0, 190] = float("nan")
    s[1, :] = 0.25
    s[2, 100:] = float("nan")
    s[3, 17] = float("inf")
    case["scores"] = s
    o = _prep(shape, _d(case))
    ref = torch.topk(s, shape.K, dim=-1)[1].sort(dim=-1)[0]
    got = o["idx32"].cpu().long()
    assert torch.equal(got[0], ref[0]) and torch.equal(got[3:], ref[3:])
    assert got[1].tolist() == list(range(shape.K))
    assert got[2].tolist() == list(range(100, 100 + shape.K))          # NaNs first, lowest index first among them
    sel = _ops().select_topk(s.to(DEV), shape.K).cpu().long()
    assert torch.equal(sel, got)
    assert int((sel < 0).sum()) == 0 and int((sel >= shape.N).sum()) == 0


def _forward_modular(shape, d, mode=None):
    """Round-1 entry points up to the similarity kernel (device tensors the new kernels are compared on)."""
    ops = _ops()
    if mode is None:
        mode = "fp32" if ops.tc_supported(shape.D, shape.K) else "fp32_fma"
    cfg = _cfg(shape, mode)
    idx = ops.select_topk(d["scores"], shape.K)
    tf = ops.addon(d["tokens"], idx, d["Wa"], d["ba"], True)
    pl = ops.prepare_prototypes(d["P"].reshape(shape.P, -1), True)
    pg = ops.prepare_prototypes(d["Pg"].reshape(shape.Pg, -1), True)
    dmin_l, argmin, act_l, dmin_g, act_g, _, _ = ops._similarity_raw(cfg, tf, pl, pg)
    return cfg, idx, tf, pl, pg, dmin_l, argmin, act_l, dmin_g, act_g


MID_CASES = [("tiny", 1), ("small", 2), ("cub_b8", 1), ("cub_b64", 3), ("cars_b64", 4), ("sweep_k196", 1)]


def _run_mid(shape, d, cfg, idx, tf, pl, dmin_l, argmin, act_l, dmin_g, act_g, train=True, use_ppc=True, upstream=1.0):
    ops, L = _ops(), _lib()
    B, K, D, P, Pg, C, m, N = shape.B, shape.K, shape.D, shape.P, shape.Pg, shape.C, shape.m, shape.N
    Bp = (B + 63) // 64 * 64
    o = dict(logits=_e(B, C), logits_g=_e(B, C), logits_l=_e(B, C), losses=torch.zeros(4, device=DEV), dlogits=_e(B, C),
             g_l=_e(B, P), g_g=_e(B, Pg), pairT=_e(P + Pg, Bp, 2), dZs_ppc=_e(B, K, D), dP_img=_e(B, m, D))
    o["ws"] = ops._ws("pph_head_mid_ws_bytes", B, K, D, P, Pg, C, m, zero=True, device=DEV)
    o["bins"] = ops._ws("pph_similarity_bwd2_ws_bytes", B, K, D, P, zero=True, device=DEV)
    for _ in range(2):      # twice: the counters must self-reset
        L.call("pph_head_mid", act_l, act_g, dmin_l, dmin_g, argmin, d["Wl"], d["Wg"], d["labels"],
               B, K, D, P, Pg, C, m, N, float(shape.global_coe), cfg.act_id, float(cfg.eps), float(upstream),
               1 if train else 0, 1 if use_ppc else 0, tf.Zs, tf.z2s, d["P"].reshape(P, -1), pl.p2, idx,
               float(cfg.ppc_cov_thresh), float(cfg.ppc_mean_thresh), 0.1, 0.5, o["ws"], o["bins"],
               o["logits"], o["logits_g"], o["logits_l"], o["losses"], o["dlogits"], o["g_l"], o["g_g"], o["pairT"],
               o["dZs_ppc"] if use_ppc else None, o["dP_img"] if use_ppc else None, None)
    torch.cuda.synchronize()
    return o


@pytest.mark.parametrize("key,seed", MID_CASES)
def test_head_mid_matches_round1_kernels_and_oracle(key, seed):
    ops, L = _ops(), _lib()
    shape = synth.SHAPES[key]
    case = synth.make_case(shape, seed=seed)
    d = _d(case)
    B, K, D, P, Pg, C, m, N = shape.B, shape.K, shape.D, shape.P, shape.Pg, shape.C, shape.m, shape.N
    cfg, idx, tf, pl, pg, dmin_l, argmin, act_l, dmin_g, act_g = _forward_modular(shape, d)
    o = _run_mid(shape, d, cfg, idx, tf, pl, dmin_l, argmin, act_l, dmin_g, act_g)
    # last layers
    lg, ll, lt = _e(B, C), _e(B, C), _e(B, C)
    L.call("pph_logits_fwd", act_l, act_g, d["Wl"], d["Wg"], B, P, Pg, C, float(shape.global_coe), lt, lg, ll)
    ref_lt = shape.global_coe * (act_g.cpu().double() @ case["Wg"].double().T) + \
        (1 - shape.global_coe) * (act_l.cpu().double() @ case["Wl"].double().T)
    assert rel_close(o["logits"].cpu(), ref_lt, 2e-5), max_rel(o["logits"].cpu(), ref_lt)
    assert rel_close(o["logits"].cpu(), lt.cpu(), 2e-5) and rel_close(o["logits_g"].cpu(), lg.cpu(), 2e-5)
    assert rel_close(o["logits_l"].cpu(), ll.cpu(), 2e-5)
    # cross-entropy + dlogits (torch on the kernel's own logits)
    lc = o["logits"].cpu().double().requires_grad_(True)
    ce = torch.nn.functional.cross_entropy(lc, case["labels"])
    ce.backward()
    assert rel_close(o["losses"][1].cpu(), ce.detach(), 1e-5)
    assert norm_rel(o["dlogits"].cpu(), lc.grad) < 1e-5
    # last-layer backward x similarity derivative
    g_l, g_g = _e(B, P), _e(B, Pg)
    L.call("pph_logits_bwd", o["dlogits"], None, None, d["Wl"], d["Wg"], dmin_l, dmin_g, B, P, Pg, C,
           float(shape.global_coe), cfg.act_id, float(cfg.eps), g_l, g_g)
    torch.cuda.synchronize()
    # the round-1 kernel contracts a 3-term bf16 split on tcgen05; this one is exact FP32 -> compare with float64
    dl = o["dlogits"].cpu().double()
    dm = dmin_l.cpu().double()
    da = torch.where(dm > 0, 1 / (dm + 1) - 1 / (dm + cfg.eps), torch.zeros_like(dm))
    ref_gl = (1 - shape.global_coe) * (dl @ case["Wl"].double()) * da
    dmg = dmin_g.cpu().double()
    dag = torch.where(dmg > 0, 1 / (dmg + 1) - 1 / (dmg + cfg.eps), torch.zeros_like(dmg))
    ref_gg = shape.global_coe * (dl @ case["Wg"].double()) * dag
    assert norm_rel(o["g_l"].cpu(), ref_gl) < 1e-5 and norm_rel(o["g_g"].cpu(), ref_gg) < 1e-5
    assert norm_rel(o["g_l"].cpu(), g_l.cpu()) < 1e-4 and norm_rel(o["g_g"].cpu(), g_g.cpu()) < 1e-4
    pt = o["pairT"].cpu()
    assert torch.equal(pt[:P, :B, 0], o["g_l"].cpu().T) and torch.equal(pt[P:, :B, 0], o["g_g"].cpu().T)
    assert torch.equal(pt[:P, :B, 1].contiguous().view(torch.int32), argmin.cpu().T.contiguous())
    assert bool((pt[P:, :B, 1].contiguous().view(torch.int32) == K).all())
    assert float(pt[:, B:].abs().max()) == 0.0 if pt.shape[1] > B else True
    # bins: a stable counting sort of every image's prototypes by token slot
    ws = o["bins"].cpu().numpy()
    stride = lambda n: (n + 255) // 256 * 256  # noqa: E731
    off_list = 2 * stride(4 * B * (K + 1))
    bs = np.frombuffer(ws[:4 * B * (K + 1)].tobytes(), dtype=np.int32).reshape(B, K + 1)
    bl = np.frombuffer(ws[off_list:off_list + 4 * B * P].tobytes(), dtype=np.int32).reshape(B, P)
    am = argmin.cpu().numpy()
    for b in range(B):
        order = np.argsort(am[b], kind="stable")
        assert np.array_equal(bl[b], order)
        assert np.array_equal(bs[b], np.searchsorted(am[b][order], np.arange(K + 1)))
    # PPC loss forward + backward against the round-1 kernels (same arithmetic) and the oracle
    cov, mean = ops.ppc_loss(cfg, tf, d["P"], pl.p2, d["labels"], m, N)
    assert rel_close(o["losses"][2].cpu(), cov.cpu(), 1e-6) and rel_close(o["losses"][3].cpu(), mean.cpu(), 1e-6)
    tot = o["losses"][1] + 0.1 * o["losses"][2] + 0.5 * o["losses"][3]
    assert rel_close(o["losses"][0].cpu(), tot.cpu(), 1e-6)
    dsl, st = _e(B, m, K), _e(B, m, 8)
    part, cnt, los = _e(B, 2), torch.zeros(1, dtype=torch.int32, device=DEV), _e(2)
    Pl2 = d["P"].reshape(P, -1).contiguous()
    L.call("pph_ppc_fwd", tf.Zs, tf.z2s, Pl2, pl.p2, idx, d["labels"], B, K, D, P, m, N, cfg.act_id, float(cfg.eps),
           float(cfg.ppc_cov_thresh), float(cfg.ppc_mean_thresh), dsl, st, part, cnt, los)
    dZp, dPp = _e(B, K, D), torch.zeros(P, D, device=DEV)
    L.call("pph_ppc_bwd", tf.Zs, Pl2, idx, d["labels"], dsl, st, None, 0.1, 0.5, B, K, D, P, m, N, cfg.act_id,
           float(cfg.eps), float(cfg.ppc_cov_thresh), float(cfg.ppc_mean_thresh), 0, dZp, dPp)
    torch.cuda.synchronize()
    assert norm_rel(o["dZs_ppc"].cpu(), dZp.cpu()) < 1e-6
    dP_sum = torch.zeros(P, D)
    lab = case["labels"].clamp(0, P // m - 1)
    for b in range(B):
        dP_sum[lab[b] * m:(lab[b] + 1) * m] += o["dP_img"][b].cpu()
    assert norm_rel(dP_sum, dPp.cpu()) < 1e-5


def test_head_mid_eval_and_no_ppc():
    shape = synth.SHAPES["cub_b8"]
    case = synth.make_case(shape, seed=5)
    d = _d(case)
    cfg, idx, tf, pl, pg, dmin_l, argmin, act_l, dmin_g, act_g = _forward_modular(shape, d)
    full = _run_mid(shape, d, cfg, idx, tf, pl, dmin_l, argmin, act_l, dmin_g, act_g)
    ev = _run_mid(shape, d, cfg, idx, tf, pl, dmin_l, argmin, act_l, dmin_g, act_g, train=False, use_ppc=False)
    assert torch.equal(ev["logits"], full["logits"]) and torch.equal(ev["losses"][1], full["losses"][1])
    assert float(ev["losses"][2]) == 0.0 and torch.equal(ev["losses"][0], ev["losses"][1])
    nop = _run_mid(shape, d, cfg, idx, tf, pl, dmin_l, argmin, act_l, dmin_g, act_g, train=True, use_ppc=False, upstream=2.0)
    assert norm_rel(nop["g_l"].cpu(), 2.0 * full["g_l"].cpu()) < 1e-6


@pytest.mark.parametrize("key,seed", MID_CASES + [("dogs_b256", 1), ("sweep_k49", 2)])
def test_similarity_bwd2_matches_round1_kernel(key, seed):
    ops, L = _ops(), _lib()
    shape = synth.SHAPES[key]
    if key == "dogs_b256":
        shape = shape.with_batch(11)
    assert L.load().pph_similarity_bwd2_supported(shape.B, shape.K, shape.D, shape.P, shape.Pg)
    case = synth.make_case(shape, seed=seed)
    d = _d(case)
    B, K, D, P, Pg, m = shape.B, shape.K, shape.D, shape.P, shape.Pg, shape.m
    cfg, idx, tf, pl, pg, dmin_l, argmin, act_l, dmin_g, act_g = _forward_modular(shape, d)
    o = _run_mid(shape, d, cfg, idx, tf, pl, dmin_l, argmin, act_l, dmin_g, act_g)
    Pl2, Pg2 = d["P"].reshape(P, -1).contiguous(), d["Pg"].reshape(Pg, -1).contiguous()
    # round-1 kernel: similarity gradients + PPC contributions through its add_ inputs
    dP_ppc = torch.zeros(P, D, device=DEV)
    lab = d["labels"].clamp(0, P // m - 1)
    for b in range(B):
        dP_ppc[lab[b] * m:(lab[b] + 1) * m] += o["dP_img"][b]
    ws = ops.bwd_workspace(B, K, D, P, Pg, DEV)
    r = dict(dZs=_e(B, K, D), dZc=_e(B, D), dP=_e(P, D), dPg=_e(Pg, D))
    L.call("pph_similarity_bwd", o["g_l"], o["g_g"], argmin, tf.Zs, tf.Zc, Pl2, Pg2, B, K, D, P, Pg, ws, 3,
           o["dZs_ppc"], dP_ppc, r["dZs"], r["dZc"], r["dP"], r["dPg"])
    n = dict(dZs=_e(B, K, D), dZc=_e(B, D), dP=_e(P, D), dPg=_e(Pg, D))

    def bwd2(parts, add, dpi, dpre_out=0):
        L.call("pph_similarity_bwd2", parts, o["g_l"], o["g_g"], o["pairT"], o["bins"], tf.Zs, tf.Zc, Pl2, Pg2, B, K, D, P,
               Pg, m, add, dpi, dpre_out, n["dZs"], n["dZc"], n["dP"], n["dPg"])

    for _ in range(2):
        bwd2(7, o["dZs_ppc"], o["dP_img"])
    torch.cuda.synchronize()
    for k in r:
        assert norm_rel(n[k].cpu(), r[k].cpu()) < 1e-5, (k, norm_rel(n[k].cpu(), r[k].cpu()))
    # the three kinds one by one (as the step launches them), with the Z (1 - Z) factor folded in
    for k in n:
        n[k].fill_(7.0)
    for parts in (1, 2, 4):
        bwd2(parts, o["dZs_ppc"], o["dP_img"], 1)
    torch.cuda.synchronize()
    assert norm_rel(n["dZs"].cpu(), (r["dZs"] * tf.Zs * (1 - tf.Zs)).cpu()) < 1e-5
    assert norm_rel(n["dZc"].cpu(), (r["dZc"] * tf.Zc * (1 - tf.Zc)).cpu()) < 1e-5
    assert norm_rel(n["dP"].cpu(), r["dP"].cpu()) < 1e-5 and norm_rel(n["dPg"].cpu(), r["dPg"].cpu()) < 1e-5
    # without the PPC inputs
    L.call("pph_similarity_bwd", o["g_l"], o["g_g"], argmin, tf.Zs, tf.Zc, Pl2, Pg2, B, K, D, P, Pg, ws, 3,
           None, None, r["dZs"], r["dZc"], r["dP"], r["dPg"])
    bwd2(7, None, None)
    torch.cuda.synchronize()
    for k in r:
        assert norm_rel(n[k].cpu(), r[k].cpu()) < 1e-5, (k, norm_rel(n[k].cpu(), r[k].cpu()))


@pytest.mark.parametrize("key,seed", [("tiny", 1), ("small", 2), ("cub_b8", 1), ("cub_b64", 3), ("cars_b64", 4),
                                      ("sweep_k144_d384", 1)])
def test_addon_bwd2_matches_float64(key, seed):
    ops, L = _ops(), _lib()
    shape = synth.SHAPES[key]
    case = synth.make_case(shape, seed=seed)
    d = _d(case)
    B, N, Din, D, K = shape.B, shape.N, shape.Din, shape.D, shape.K
    idx = ops.select_topk(d["scores"], K)
    tf = ops.addon(d["tokens"], idx, d["Wa"], d["ba"], False)
    g = torch.Generator().manual_seed(seed)
    dZs, dZc = torch.randn(B, K, D, generator=g) * 1e-2, torch.randn(B, D, generator=g) * 1e-2
    ws = ops._ws("pph_addon_bwd2_ws_bytes", B, N, Din, D, K, zero=True, device=DEV)
    dWa, dba, dtok = _e(D, Din), _e(D), torch.full((B, 1 + N, Din), 7.0, device=DEV)
    dpre_s = (dZs.to(DEV) * tf.Zs * (1 - tf.Zs)).contiguous()
    dpre_c = (dZc.to(DEV) * tf.Zc * (1 - tf.Zc)).contiguous()
    for parts in ((3,), (3,), (1, 2)):          # both roles in one launch (twice: self-resetting tickets), then separately
        dWa.fill_(7.0)
        dba.fill_(7.0)
        dtok.fill_(7.0)
        for pt in parts:
            L.call("pph_addon_bwd2", pt, d["tokens"], idx, d["Wa"].reshape(D, Din), dpre_s, dpre_c, B, N, Din, D, K, ws,
                   dWa, dba, dtok)
    torch.cuda.synchronize()
    Z = torch.cat([tf.Zs.cpu(), tf.Zc.cpu()[:, None]], 1).double()
    dZ = torch.cat([dZs, dZc[:, None]], 1).double()
    dpre = dZ * Z * (1 - Z)                                           # (B, K+1, D)
    rows = torch.cat([1 + idx.cpu().long(), torch.zeros(B, 1, dtype=torch.long)], 1)        # source token row
    X = torch.gather(case["tokens"].double(), 1, rows[:, :, None].expand(B, K + 1, Din))
    ref_dWa = torch.einsum("bkd,bki->di", dpre, X)
    ref_dba = dpre.sum((0, 1))
    ref_dtok = torch.zeros(B, 1 + N, Din, dtype=torch.float64)
    ref_dtok.scatter_(1, rows[:, :, None].expand(B, K + 1, Din), dpre @ case["Wa"].reshape(D, Din).double())
    assert norm_rel(dWa.cpu(), ref_dWa) < 1e-5 and norm_rel(dba.cpu(), ref_dba) < 1e-5
    assert norm_rel(dtok.cpu(), ref_dtok) < 1e-5
    assert bool((dtok.cpu()[ref_dtok == 0] == 0).all())               # unselected token rows are exact zeros


TC2_CASES = [("cub_b8", 1), ("cub_b64", 3), ("cars_b64", 4), ("sweep_k49", 2), ("sweep_k196", 1)]


@pytest.mark.parametrize("key,seed", TC2_CASES)
def test_addon_fwd2_single_shot_matches_oracle(key, seed):
    """pph_addon_fwd2 (single-shot tcgen05, columns split over CTAs, norms completed by the last column tile)."""
    ops, L = _ops(), _lib()
    shape = synth.SHAPES[key]
    case = synth.make_case(shape, seed=seed)
    d = _d(case)
    B, N, Din, D, K = shape.B, shape.N, shape.Din, shape.D, shape.K
    assert L.load().pph_addon_tc2_supported(B, N, Din, D, K) & 1
    idx = ops.select_topk(d["scores"], K)
    bf = torch.bfloat16
    o = dict(Zs=_e(B, K, D), Zc=_e(B, D), z2s=_e(B, K), z2c=_e(B), z2s_ctr=_e(B, K), z2c_ctr=_e(B), z2s_hi=_e(B, K),
             z2c_hi=_e(B), Zs_hi=_e(B * K, D, dt=bf), Zs_lo=_e(B * K, D, dt=bf), Zc_hi=_e(B, D, dt=bf), Zc_lo=_e(B, D, dt=bf))
    ws = ops._ws("pph_addon_tc2_ws_bytes", B, N, Din, D, K, zero=True, device=DEV)
    for _ in range(2):
        L.call("pph_addon_fwd2", d["tokens"], idx, d["Wa"].reshape(D, Din), d["ba"], B, N, Din, D, K, o["Zs"], o["Zc"],
               o["z2s"], o["z2c"], 0.5, o["z2s_ctr"], o["z2c_ctr"], o["z2s_hi"], o["z2c_hi"], o["Zs_hi"], o["Zs_lo"],
               o["Zc_hi"], o["Zc_lo"], ws)
    torch.cuda.synchronize()
    Zs, Zc = O.addon(case["tokens"], idx.cpu().long(), case["Wa"], case["ba"])
    # 3-term bf16 split on the pre-activation (~1e-5) and sigmoid through ex2.approx / rcp.approx: 5e-5 on Z
    assert rel_close(o["Zs"].cpu(), Zs, 5e-5) and rel_close(o["Zc"].cpu(), Zc, 5e-5), max_rel(o["Zs"].cpu(), Zs)
    assert rel_close(o["z2s"].cpu(), (Zs * Zs).sum(-1), 2e-5) and rel_close(o["z2c"].cpu(), (Zc * Zc).sum(-1), 2e-5)
    assert rel_close(o["z2s"].cpu(), (o["Zs"] ** 2).sum(-1).cpu(), 2e-6)          # norms of the kernel's own output
    assert rel_close(o["z2s_ctr"].cpu(), ((o["Zs"] - 0.5) ** 2).sum(-1).cpu(), 2e-6)
    assert rel_close(o["z2c_ctr"].cpu(), ((o["Zc"] - 0.5) ** 2).sum(-1).cpu(), 2e-6)
    rec = (o["Zs_hi"].float() + o["Zs_lo"].float()).cpu().reshape(Zs.shape)
    assert float((rec - (o["Zs"].cpu() - 0.5)).abs().max()) < 1e-5
    assert rel_close(o["z2s_hi"].cpu(), (o["Zs_hi"].float() ** 2).sum(-1).reshape(B, K).cpu(), 1e-5)
    assert rel_close(o["z2c_hi"].cpu(), (o["Zc_hi"].float() ** 2).sum(-1).cpu(), 1e-5)


@pytest.mark.parametrize("key,seed,heads", [(k, s, 0) for k, s in TC2_CASES] + [("cub_b8", 5, 3), ("cub_b64", 6, 6)])
def test_fused_selection_addon_forward_equals_the_two_launches(key, seed, heads):
    """pph_select_addon_fwd = pph_select_topk (bit-exact index lists, ties and NaN included) + pph_addon_fwd2 (bitwise the
    same features, operands and norms: same arithmetic, the row list just comes from shared memory)."""
    ops, L = _ops(), _lib()
    shape = synth.SHAPES[key]
    case = synth.make_case(shape, seed=seed, heads=heads)
    d = _d(case)
    B, N, Din, D, K = shape.B, shape.N, shape.Din, shape.D, shape.K
    assert L.load().pph_addon_tc2_supported(B, N, Din, D, K) & 8
    scores = d["scores_h"] if heads else d["scores"].clone()
    if not heads:                      # ties and a NaN: the ranking rule must be the select kernel's
        scores[0, 5] = scores[0, 9]
        scores[1 % B, 7] = float("nan")
        scores[B - 1, :4] = scores[B - 1, 4]
    idx = ops.select_topk(scores, K)                 # (held to the oracle / torch.topk, ties and NaN, in test_gpu_parity.py)
    if heads:
        assert torch.equal(idx.cpu().long(), O.select_tokens(scores.cpu(), K))
    bf = torch.bfloat16

    def outs():
        return dict(Zs=_e(B, K, D), Zc=_e(B, D), z2s=_e(B, K), z2c=_e(B), z2s_ctr=_e(B, K), z2c_ctr=_e(B), z2s_hi=_e(B, K),
                    z2c_hi=_e(B), Zs_hi=_e(B * K, D, dt=bf), Zs_lo=_e(B * K, D, dt=bf), Zc_hi=_e(B, D, dt=bf),
                    Zc_lo=_e(B, D, dt=bf))
    a, b = outs(), outs()
    ws = ops._ws("pph_addon_tc2_ws_bytes", B, N, Din, D, K, zero=True, device=DEV)
    L.call("pph_addon_fwd2", d["tokens"], idx, d["Wa"].reshape(D, Din), d["ba"], B, N, Din, D, K, a["Zs"], a["Zc"],
           a["z2s"], a["z2c"], 0.5, a["z2s_ctr"], a["z2c_ctr"], a["z2s_hi"], a["z2c_hi"], a["Zs_hi"], a["Zs_lo"],
           a["Zc_hi"], a["Zc_lo"], ws)
    idx2 = torch.full((B, K), -1, dtype=torch.int32, device=DEV)
    for _ in range(2):
        L.call("pph_select_addon_fwd", scores, max(heads, 1), d["tokens"], d["Wa"].reshape(D, Din), d["ba"], B, N, Din, D, K,
               idx2, b["Zs"], b["Zc"], b["z2s"], b["z2c"], 0.5, b["z2s_ctr"], b["z2c_ctr"], b["z2s_hi"], b["z2c_hi"],
               b["Zs_hi"], b["Zs_lo"], b["Zc_hi"], b["Zc_lo"], ws)
    torch.cuda.synchronize()
    assert torch.equal(idx2, idx)
    for k in a:
        assert torch.equal(a[k].view(torch.int16 if a[k].dtype == bf else torch.int32),
                           b[k].view(torch.int16 if b[k].dtype == bf else torch.int32)), k


@pytest.mark.parametrize("key,seed", TC2_CASES)
def test_addon_bwd3_single_shot_matches_float64(key, seed):
    ops, L = _ops(), _lib()
    shape = synth.SHAPES[key]
    case = synth.make_case(shape, seed=seed)
    d = _d(case)
    B, N, Din, D, K = shape.B, shape.N, shape.Din, shape.D, shape.K
    bits = L.load().pph_addon_tc2_supported(B, N, Din, D, K)
    assert bits & 2
    idx = ops.select_topk(d["scores"], K)
    g = torch.Generator().manual_seed(seed)
    dpre_s = (torch.randn(B, K, D, generator=g) * 1e-3).to(DEV)
    dpre_c = (torch.randn(B, D, generator=g) * 1e-3).to(DEV)
    ws = ops._ws("pph_addon_tc2_ws_bytes", B, N, Din, D, K, zero=True, device=DEV)
    dWa, dba, dtok = _e(D, Din), _e(D), torch.zeros(B, 1 + N, Din, device=DEV)
    parts = 3 if bits & 4 else 2
    half = (0.5 * dpre_s).contiguous()          # dpre_s = half + half through the dpre_add_s operand
    for add in (None, half):
        L.call("pph_addon_bwd3", parts, d["tokens"], idx, d["Wa"].reshape(D, Din), dpre_s if add is None else half, dpre_c,
               add, B, N, Din, D, K, ws, dWa, dba, dtok)
    torch.cuda.synchronize()
    dpre = torch.cat([dpre_s.cpu(), dpre_c.cpu()[:, None]], 1).double()
    rows = torch.cat([1 + idx.cpu().long(), torch.zeros(B, 1, dtype=torch.long)], 1)
    X = torch.gather(case["tokens"].double(), 1, rows[:, :, None].expand(B, K + 1, Din))
    ref_dtok = torch.zeros(B, 1 + N, Din, dtype=torch.float64)
    ref_dtok.scatter_(1, rows[:, :, None].expand(B, K + 1, Din), dpre @ case["Wa"].reshape(D, Din).double())
    assert norm_rel(dtok.cpu(), ref_dtok) < 2e-5, norm_rel(dtok.cpu(), ref_dtok)
    if bits & 4:
        assert norm_rel(dWa.cpu(), torch.einsum("bkd,bki->di", dpre, X)) < 2e-5
        assert norm_rel(dba.cpu(), dpre.sum((0, 1))) < 2e-5


_S = dict(prep="simt", addon_bwd="simt", bwd="staged", ppc="inline", select="kernel")
_T = dict(prep="tc", addon_bwd="tc", bwd="gather", ppc="late", select="fused")


@pytest.mark.parametrize("variants", [_S, _T, dict(_T, ppc="inline"), dict(_T, ppc="split"), dict(_T, bwd="staged", ppc="split"), dict(_S, bwd="gather"),
                                      dict(_S, ppc="split"), dict(_T, prep="simt"), dict(_T, addon_bwd="simt"), dict(_T, select="kernel")])
def test_step_variants_agree_with_the_oracle(variants):
    shape, case, g, fn = load_golden("cub_b8_s1")
    step, params = _make_step(shape, case, "fp32", variants=variants)
    assert step.fused.variants == variants
    step.run(0)
    step.run(0)
    torch.cuda.synchronize()
    f = step.fused
    assert np.array_equal(f.idx32.cpu().numpy(), g["idx"])
    assert rel_close(f.logits.cpu(), g["logits_train"], 1e-4)
    assert rel_close(f.losses.cpu()[0], g["loss"], 1e-4)
    ref = O.head_train_step(case, shape, fn=fn, route=f.argmin.cpu().long())
    got = dict(g_tokens=f.dtokens, g_P=params["P"].grad, g_Pg=params["Pg"].grad, g_Wa=params["Wa"].grad,
               g_ba=params["ba"].grad)
    for k, v in got.items():
        e = norm_rel(v.cpu().reshape(ref[k].shape), ref[k])
        assert e < 2e-4, (variants, k, e)


# ---------------------------------------------------------------------------------------------------------------
# whole step
# ---------------------------------------------------------------------------------------------------------------
def _make_step(shape, case, mode, impl="auto", train=True, use_ppc=True, variants=None, fn="log"):
    from protopformer_b200.graph import GraphedHeadStep
    params = {k: case[k].to(DEV).clone() for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
    for k in ("Wa", "ba", "P", "Pg"):
        params[k].requires_grad_(train)
    step = GraphedHeadStep(params, _cfg(shape, mode, fn), B=shape.B, N=shape.N, C=shape.C, m=shape.m, train=train, impl=impl,
                           use_ppc=use_ppc, variants=variants)
    step.load(0, case["tokens"], case["scores"], case["labels"])
    torch.cuda.synchronize()
    step.capture()
    return step, params


ALL_GOLDEN = [(n, "fp32") for n in GOLDEN_CASES if n not in ("tiny_s1", "tiny_s2_linear", "small_s1", "small_s3_matched")] + \
    [("cub_b8_s1", "bf16"), ("small_s1", "fp32_fma"), ("tiny_s1", "fp32_fma"), ("small_s3_matched", "fp32_fma"),
     ("tiny_s2_linear", "fp32_fma")]


@pytest.mark.parametrize("name,mode", ALL_GOLDEN)
def test_five_launch_step_matches_reference_fixture(name, mode):
    shape, case, g, fn = load_golden(name)
    step, params = _make_step(shape, case, mode, fn=fn)
    v2 = _ops().fused_step_supported(shape.B, shape.N, shape.Din, shape.D, shape.K, shape.P, shape.Pg, shape.C, shape.m)
    assert step.impl == ("v2" if v2 else "v1") and (not v2 or step.kernel_launches_per_step <= 12)
    step.run(0)
    step.run(0)
    torch.cuda.synchronize()
    f = step.fused
    tol = 5e-3 if mode == "bf16" else (1e-3 if "matched" in name else 1e-4)
    assert np.array_equal(f.idx32.cpu().numpy(), g["idx"])
    assert rel_close(f.logits.cpu(), g["logits_train"], tol), max_rel(f.logits.cpu(), g["logits_train"])
    losses = f.losses.cpu()
    assert rel_close(losses[0], g["loss"], tol) and rel_close(losses[1], g["ce"], tol)
    ptol = 1e-3 if "matched" in name else 1e-4
    assert rel_close(losses[2], g["ppc_cov"], ptol) and rel_close(losses[3], g["ppc_mean"], ptol)
    ref = O.head_train_step(case, shape, fn=fn, route=f.argmin.cpu().long())
    gt = 5e-2 if mode == "bf16" else (5e-3 if "matched" in name else 2e-4)
    got = dict(g_tokens=f.dtokens, g_P=params["P"].grad, g_Pg=params["Pg"].grad, g_Wa=params["Wa"].grad,
               g_ba=params["ba"].grad)
    for k, v in got.items():
        e = norm_rel(v.cpu().reshape(ref[k].shape), ref[k])
        assert e < gt, (k, e)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_benchmarked_shape_against_oracle(mode):
    """the shape bench.py times (cub_b64, B = 64) compared with the CPU oracle itself -- every output
    and every gradient, both precision modes -- not with another CUDA path."""
    shape = synth.SHAPES["cub_b64"]
    case = synth.make_case(shape, seed=1)              # the batch bench.py's parameters come from
    step, params = _make_step(shape, case, mode)
    step.run(0)
    torch.cuda.synchronize()
    f = step.fused
    ref = O.head_train_step(case, shape, route=f.argmin.cpu().long())
    free = O.head_train_step(case, shape)               # oracle's own routing: argmin comparison
    assert torch.equal(f.idx32.cpu().long(), ref["idx"])
    flips = float((f.argmin.cpu().long() != free["argmax"]).float().mean())
    tol = dict(fp32=dict(act=1e-4, logits=1e-4, loss=1e-4, grad=1e-4, flips=1e-4),
               bf16=dict(act=4e-3, logits=5e-4, loss=5e-4, grad=2e-2, flips=0.01))[mode]
    assert flips <= tol["flips"], flips
    assert rel_close(f.act_l.cpu(), ref["act_l"], tol["act"]), max_rel(f.act_l.cpu(), ref["act_l"])
    assert rel_close(f.act_g.cpu(), ref["act_g"], tol["act"])
    assert rel_close(f.dmin_l.cpu(), ref["dmin_l"], tol["act"])
    assert rel_close(f.logits.cpu(), ref["logits"], tol["logits"]), max_rel(f.logits.cpu(), ref["logits"])
    losses = f.losses.cpu()
    assert rel_close(losses[0], ref["loss"], tol["loss"]) and rel_close(losses[1], ref["ce"], tol["loss"])
    assert rel_close(losses[2], ref["ppc_cov"], 1e-4) and rel_close(losses[3], ref["ppc_mean"], 1e-4)
    got = dict(g_tokens=f.dtokens, g_P=params["P"].grad, g_Pg=params["Pg"].grad, g_Wa=params["Wa"].grad,
               g_ba=params["ba"].grad)
    errs = {k: norm_rel(v.cpu().reshape(ref[k].shape), ref[k]) for k, v in got.items()}
    print("cub_b64", mode, "argmin flips", flips, "gradient errors (max |a-b| / max |b|)", errs)
    for k, e in errs.items():
        assert e < tol["grad"], (k, e)


def test_step_is_bit_reproducible_and_matches_round1_sequence():
    shape = synth.SHAPES["cub_b64"]
    case = synth.make_case(shape, seed=4)
    s2, p2 = _make_step(shape, case, "fp32")
    s2.run(0)
    torch.cuda.synchronize()
    a = {k: p2[k].grad.clone() for k in ("P", "Pg", "Wa", "ba")}
    la, dta = s2.fused.losses.clone(), s2.fused.dtokens.clone()
    for _ in range(3):
        s2.run(0)
    torch.cuda.synchronize()
    for k in a:                                          # no atomics on data anywhere: bitwise equal replays
        assert torch.equal(a[k], p2[k].grad), k
    assert torch.equal(la, s2.fused.losses) and torch.equal(dta, s2.fused.dtokens)
    s1, p1 = _make_step(shape, case, "fp32", impl="v1")
    s1.run(0)
    torch.cuda.synchronize()
    assert rel_close(s1.fused.losses.cpu(), la.cpu(), 1e-4)
    # the round-1 sequence is the less accurate arm (tcgen05 3-term split in the small GEMMs, routed through slightly
    # different Z): the round-2 step is held to the oracle directly (test_benchmarked_shape_against_oracle)
    for k in ("P", "Pg", "Wa", "ba"):
        assert norm_rel(p1[k].grad.cpu(), a[k].cpu()) < 2e-4, (k, norm_rel(p1[k].grad.cpu(), a[k].cpu()))
    assert norm_rel(s1.fused.dtokens.cpu(), dta.cpu()) < 2e-4


def test_five_launch_eval_step_and_no_ppc():
    shape, case, g, fn = load_golden("cub_b8_s1")
    step, _ = _make_step(shape, case, "fp32", train=False)
    step.run(0)
    torch.cuda.synchronize()
    assert step.kernel_launches_per_step <= 6
    assert rel_close(step.fused.logits.cpu(), g["logits"], 1e-4)
    s_np, p_np = _make_step(shape, case, "fp32", use_ppc=False)
    s_np.run(0)
    torch.cuda.synchronize()
    ref = O.head_train_step(case, shape, ppc_cov_coe=0.0, ppc_mean_coe=0.0, route=s_np.fused.argmin.cpu().long())
    assert rel_close(s_np.fused.losses[0].cpu(), ref["ce"], 1e-4)
    assert norm_rel(p_np["P"].grad.cpu().reshape(ref["g_P"].shape), ref["g_P"]) < 1e-4


@pytest.mark.parametrize("B", [1, 65, 256])
def test_five_launch_step_other_batch_sizes(B):
    shape = synth.SHAPES["cub_b64"].with_batch(B)
    case = synth.make_case(shape, seed=B)
    s2, p2 = _make_step(shape, case, "fp32")
    s1, p1 = _make_step(shape, case, "fp32", impl="v1")
    for s in (s1, s2):
        s.run(0)
    torch.cuda.synchronize()
    assert torch.equal(s1.fused.idx32, s2.fused.idx32)
    # the two add-on kernels differ by ~1e-6 on Z (tcgen05 3-term split vs exact FP32): near-tie argmins may move
    assert float((s1.fused.argmin != s2.fused.argmin).float().mean()) < 1e-3
    assert rel_close(s1.fused.losses.cpu(), s2.fused.losses.cpu(), 1e-4)
    assert rel_close(s1.fused.logits.cpu(), s2.fused.logits.cpu(), 1e-4)
    for k in ("P", "Pg", "Wa", "ba"):
        assert norm_rel(p1[k].grad.cpu(), p2[k].grad.cpu()) < 5e-3, k
    assert norm_rel(s1.fused.dtokens.cpu(), s2.fused.dtokens.cpu()) < 5e-3


def test_selection_first_host_transfer_feeds_the_same_step():
    """GraphedHeadStep.load_host: only the CLS row and the selected rows cross the bus (read by a kernel from pinned host
    memory); the step's results are bitwise those of a full copy of the batch."""
    shape = synth.SHAPES["cub_b8"]
    case = synth.make_case(shape, seed=11)
    step, params = _make_step(shape, case, "fp32")
    step.run(0)
    torch.cuda.synchronize()
    want = (step.fused.losses.clone(), step.fused.dtokens.clone(), params["P"].grad.clone(), params["Wa"].grad.clone())
    idx = step.fused.idx32.long().cpu()
    step.tokens[0].detach().fill_(float("nan"))                       # stale rows must never be read
    host = {k: case[k].pin_memory() for k in ("tokens", "scores", "labels")}
    moved = step.load_host(0, host["tokens"], host["scores"], host["labels"])
    assert moved == 4 * shape.B * shape.N + 8 * shape.B + 4 * shape.B * (shape.K + 1) * shape.Din
    torch.cuda.synchronize()
    dev_tok = step.tokens[0].detach().cpu()
    for b in range(shape.B):
        rows = torch.cat([torch.zeros(1, dtype=torch.long), idx[b] + 1])
        assert torch.equal(dev_tok[b, rows], case["tokens"][b, rows])
        other = torch.ones(shape.N + 1, dtype=torch.bool)
        other[rows] = False
        assert torch.isnan(dev_tok[b, other]).all()                   # nothing else was transferred
    step.run(0)
    torch.cuda.synchronize()
    got = (step.fused.losses, step.fused.dtokens, params["P"].grad, params["Wa"].grad)
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    from protopformer_b200 import _lib as L
    # the host pipeline graphs: selection taken from the transfer, loss mirrored into pinned host memory by the kernel
    step.capture_host_pipeline()
    step.load_host(0, host["tokens"], host["scores"], host["labels"])
    lh = step.run_host(0)
    torch.cuda.synchronize()
    assert torch.equal(lh, want[0].cpu())
    for a, b in zip(want[1:], (step.fused.dtokens, params["P"].grad, params["Wa"].grad)):
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError):                                  # pageable memory is refused, not copied silently
        L.call("pph_gather_rows_host", case["tokens"].data_ptr(), step._load_idx[0], shape.B, shape.N, shape.Din, shape.K,
               step.tokens[0].detach(), 8)
