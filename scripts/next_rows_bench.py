"""Device time of the remaining 'next' rows: class-row maps (next #4), CaiT rollout with its start row, fused top-K."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protopformer_b200 import synth
from protopformer_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


with torch.no_grad():
    for B in (64, 256):
        shape = synth.SHAPES["cub_b64"].with_batch(B)
        case = {k: v.to(dev) for k, v in synth.make_case(shape, seed=1).items()}
        cfg = ops.HeadConfig(K=shape.K, mode="fp32_fma")
        idx = ops.select_topk(case["scores"], shape.K)
        tf = ops.addon(case["tokens"], idx, case["Wa"], case["ba"], False)
        p2 = ops.prepare_prototypes(case["P"], False).p2
        us = timeit(lambda: ops.class_activation_maps(cfg, tf, case["P"], case["labels"], shape.m, shape.N, p2l=p2))
        us_full = timeit(lambda: ops.materialize_maps(cfg, tf, case["P"], case["Pg"]), reps=5)
        print(json.dumps(dict(row="class_maps", B=B, us=round(us, 1), out_kb_per_image=shape.m * shape.N * 4 / 1e3,
                              materialise_full_map_us=round(us_full, 1))), flush=True)
    g = torch.Generator(device=dev).manual_seed(0)
    B, H, T = 64, 4, 196
    patch = [torch.softmax(2.0 * torch.randn(B, H, T, T, device=dev, generator=g), dim=-1) for _ in range(24)]
    cls = [torch.softmax(2.0 * torch.randn(B, H, 1, T + 1, device=dev, generator=g), dim=-1) for _ in range(2)]
    us = timeit(lambda: ops.rollout_scores_cait(patch + cls, 24), reps=5)
    nbytes = sum(a.numel() * 4 for a in patch + cls)
    print(json.dumps(dict(row="rollout_cait_xxs24", B=B, us=round(us, 1), gbs=round(nbytes / us / 1e3, 1))), flush=True)
    del patch, cls
    attn = [torch.softmax(2.0 * torch.randn(64, 3, 197, 197, device=dev, generator=g), dim=-1) for _ in range(11)]
    a = timeit(lambda: ops.rollout_scores(attn), reps=10)
    b = timeit(lambda: ops.rollout_scores(attn, topk=81), reps=10)
    c = timeit(lambda: ops.select_topk(ops.rollout_scores(attn), 81), reps=10)
    print(json.dumps(dict(row="rollout_deit_tiny", B=64, scores_only_us=round(a, 1), fused_topk_us=round(b, 1),
                          separate_select_us=round(c, 1))), flush=True)
