"""Measured errors of the modular CUDA path against the reference fixtures / the oracle, per precision mode, over every
golden case: the numbers the asserted tolerances in tests/test_gpu_parity.py are derived from (2x the measured maximum)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import protohead_oracle as O  # noqa: E402
from tests import test_gpu_parity as T  # noqa: E402
from tests.util import GOLDEN_CASES, load_golden, max_rel, norm_rel  # noqa: E402
from protopformer_b200 import synth  # noqa: E402

worst = {}


def upd(mode, group, key, val):
    d = worst.setdefault(mode, {}).setdefault(group, {})
    d[key] = max(d.get(key, 0.0), float(val))


for name in GOLDEN_CASES:
    shape, case, g, fn = load_golden(name)
    grp = "matched" if "matched" in name else "init"
    for mode in T._modes_for_static(shape):
        out, leaves, d = T._forward(shape, case, mode, fn, grad=True)
        for k, v in dict(logits=out.logits, act_l=out.act_l, dmin_l=out.dmin_l).items():
            upd(mode, grp, k, max_rel(v.detach().cpu(), g[k]))
        upd(mode, grp, "argmin_flip_frac", (out.argmin.cpu().long() != torch.as_tensor(g["argmax"]).long()).float().mean())
        cov, mean = T._ops().ppc_loss(T._cfg(shape, mode, fn), out.tf, leaves["P"], out.p2l, d["labels"], shape.m, shape.N)
        loss = F.cross_entropy(out.logits, d["labels"]) + 0.1 * cov + 0.5 * mean
        loss.backward()
        upd(mode, grp, "loss", abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])))
        ref = O.head_train_step(case, shape, fn=fn, route=out.argmin.cpu().long())
        for k, t in dict(g_tokens=leaves["tokens"].grad, g_P=leaves["P"].grad, g_Pg=leaves["Pg"].grad, g_Wa=leaves["Wa"].grad,
                         g_ba=leaves["ba"].grad).items():
            upd(mode, grp, k, norm_rel(t.cpu(), ref[k]))
print(json.dumps(worst, indent=1))

# where does the gradient error come from?  per case, fp32_fma mode: CUDA vs fp32 oracle, CUDA vs float64 oracle, and the
# fp32 oracle itself vs float64
print("# per case (mode fp32_fma / fp32 where available): g_tokens, g_P, g_Wa as (cuda-vs-f32 oracle | cuda-vs-f64 | f32 oracle-vs-f64)")
for name in GOLDEN_CASES:
    shape, case, g, fn = load_golden(name)
    for mode in T._modes_for_static(shape)[:2]:
        out, leaves, d = T._forward(shape, case, mode, fn, grad=True)
        cov, mean = T._ops().ppc_loss(T._cfg(shape, mode, fn), out.tf, leaves["P"], out.p2l, d["labels"], shape.m, shape.N)
        (F.cross_entropy(out.logits, d["labels"]) + 0.1 * cov + 0.5 * mean).backward()
        route = out.argmin.cpu().long()
        r32 = O.head_train_step(case, shape, fn=fn, route=route)
        r64 = O.head_train_step(case, shape, fn=fn, route=route, dtype=torch.float64)
        row = []
        for k, t in dict(g_tokens=leaves["tokens"].grad, g_P=leaves["P"].grad, g_Wa=leaves["Wa"].grad).items():
            row.append(f"{k} {norm_rel(t.cpu(), r32[k]):.1e}|{norm_rel(t.cpu().double(), r64[k]):.1e}|{norm_rel(r32[k].double(), r64[k]):.1e}")
        print(f"{name:22s} {mode:8s} " + "  ".join(row))

# localise the token-gradient error of the modular path on cub_b8_s1: CLS rows vs patch rows, and against the graphed step
shape, case, g, fn = load_golden("cub_b8_s1")
out, leaves, d = T._forward(shape, case, "fp32", fn, grad=True)
cov, mean = T._ops().ppc_loss(T._cfg(shape, "fp32", fn), out.tf, leaves["P"], out.p2l, d["labels"], shape.m, shape.N)
ce = F.cross_entropy(out.logits, d["labels"])
for label, loss in (("ce only", ce), ("ce + ppc", ce + 0.1 * cov + 0.5 * mean)):
    for v in leaves.values():
        v.grad = None
    loss.backward(retain_graph=True)
    route = out.argmin.cpu().long()
    r64 = O.head_train_step(case, shape, fn=fn, route=route, dtype=torch.float64,
                            ppc_cov_coe=0.0 if label == "ce only" else 0.1, ppc_mean_coe=0.0 if label == "ce only" else 0.5)
    gt, rt = leaves["tokens"].grad.cpu().double(), r64["g_tokens"]
    den = rt.abs().max()
    print(f"# {label}: CLS rows {float((gt[:, 0] - rt[:, 0]).abs().max() / den):.1e}  patch rows {float((gt[:, 1:] - rt[:, 1:]).abs().max() / den):.1e}"
          f"  (max |g| CLS {float(rt[:, 0].abs().max()):.2e}, patch {float(rt[:, 1:].abs().max()):.2e})")
