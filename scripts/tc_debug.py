"""GPU debugging aid: tcgen05 similarity kernel vs the FP32-FMA kernel on a few shapes (prints error statistics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protopformer_b200 import synth
from protopformer_b200 import ops, _lib

dev = torch.device("cuda:0")
shapes = [("cub_b8", None), ("cub_b64", None), ("cars_b64", 5), ("dogs_b256", 7)]
extra = [synth.HeadShape("k100", 4, 196, 192, 192, 100, 1000, 500, 100), synth.HeadShape("d64", 3, 196, 64, 64, 49, 256, 128, 16)]
allshapes = []
for k, b in shapes:
    s = synth.SHAPES[k]
    allshapes.append(s.with_batch(b) if b else s)
allshapes += extra
for s in allshapes:
    case = {k: v.to(dev) for k, v in synth.make_case(s, seed=3).items()}
    res = {}
    for mode in ("fp32_fma", "fp32", "bf16"):
        cfg = ops.HeadConfig(K=s.K, global_coe=s.global_coe, mode=mode)
        try:
            with torch.no_grad():
                o = ops.head_forward(cfg, case["tokens"], case["scores"], case["Wa"], case["ba"], case["P"], case["Pg"], case["Wl"], case["Wg"])
            torch.cuda.synchronize()
            res[mode] = o
        except Exception as e:
            print(f"[{s.name} B={s.B}] mode {mode} FAILED: {e}")
            raise
    a = res["fp32_fma"]
    for mode in ("fp32", "bf16"):
        b = res[mode]
        def mr(x, y):
            return float(((x - y).abs() / (y.abs() + 1e-6)).max())
        mism = int((a.argmin != b.argmin).sum())
        print(f"[{s.name} B={s.B} K={s.K} D={s.D} P={s.P}] {mode}: dmin_l {mr(b.dmin_l, a.dmin_l):.2e} act_l {mr(b.act_l, a.act_l):.2e} "
              f"dmin_g {mr(b.dmin_g, a.dmin_g):.2e} logits {mr(b.logits, a.logits):.2e} argmin mismatches {mism}/{a.argmin.numel()}")
        if mode == "fp32" and mr(b.dmin_l, a.dmin_l) > 1e-3:
            print("  sample ref :", a.dmin_l[0, :6].tolist(), a.argmin[0, :6].tolist())
            print("  sample tc  :", b.dmin_l[0, :6].tolist(), b.argmin[0, :6].tolist())
            print("  sample ref last rows:", a.dmin_l[-1, -4:].tolist(), " tc:", b.dmin_l[-1, -4:].tolist())
print("tc_debug done; launches:", _lib.launch_count())
