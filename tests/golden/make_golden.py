"""Generate golden fixtures from the UNMODIFIED reference (run in the build container, where /root/reference exists).

    python tests/golden/make_golden.py

For every (shape, seed, proto_mode, activation) case below the reference's own PPNet (protopformer.py) is run on
CPU fp32 through oracle/ref_harness.py, and its outputs are stored in tests/golden/<case>.npz:
  * small tensors in full (idx, argmax, logits, pooled activations, min distances, losses);
  * big tensors (gradients, maps) in full for the tiny/small shapes and as strided samples + row norms for the
    BASELINE shapes, so the fixtures stay small;
  * `near_tie`: the (b,p) pairs whose top-2 distance gap is < 1e-4 in the reference (argmax is only compared
    elsewhere);
  * `chk_*`: fingerprints of the synthetic inputs, so the tests notice if the input generator ever drifts.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from protopformer_b200 import synth  # noqa: E402

# name, shape key, batch override, seed, proto_mode, activation
CASES = [
    ("tiny_s1", "tiny", None, 1, "init", "log"),
    ("tiny_s2_linear", "tiny", None, 2, "init", "linear"),
    ("small_s1", "small", None, 1, "init", "log"),
    ("small_s3_matched", "small", None, 3, "matched", "log"),
    ("cub_b8_s1", "cub_b8", None, 1, "init", "log"),
    ("cub_b8_s2_matched", "cub_b8", None, 2, "matched", "log"),
    ("cars_b4_s1", "cars_b64", 4, 1, "init", "log"),
    ("dogs_b4_s1", "dogs_b256", 4, 1, "init", "log"),
    # sweep corners (tests/util.EXTRA_GOLDEN_CASES): every token of the image kept (K = N), few tokens, wide features
    ("sweep_k49_s1", "sweep_k49", None, 1, "init", "log"),
    ("sweep_k196_s1", "sweep_k196", None, 1, "init", "log"),
    ("sweep_k144_d384_s1", "sweep_k144_d384", None, 1, "init", "log"),
]
FULL = {"tiny", "small"}
ROW_STRIDE = 16


def pack(name, shape, seed, mode, fn):
    torch.manual_seed(0)
    torch.set_num_threads(1)       # deterministic reduction order in the fixture
    case = synth.make_case(shape, seed=seed, proto_mode=mode)
    ref = ref_harness.run_reference(case, shape, fn=fn)
    out = {}
    for k in ("idx", "argmax"):
        out[k] = ref[k].numpy().astype(np.int32)
    for k in ("logits", "logits_global", "logits_local", "logits_train", "act_l", "dmin_l", "g_ba"):
        out[k] = ref[k].numpy()
    for k in ("ce", "ppc_cov", "ppc_mean", "loss"):
        out[k] = np.float32(ref[k].item())
    d = ref["dist_map"]
    top2 = d.topk(2, dim=-1, largest=False).values
    out["near_tie"] = (top2[..., 1] - top2[..., 0] < 1e-4).nonzero().numpy().astype(np.int32)
    full = shape.name in FULL
    for k in ("g_tokens", "g_P", "g_Pg", "g_Wa"):
        t = ref[k].reshape(-1, ref[k].shape[-1])
        out[k + "_rownorm"] = t.norm(dim=-1).numpy()
        out[k] = t.numpy() if full else t[::ROW_STRIDE].numpy()
    for k in ("act_map", "dist_map"):
        t = ref[k]
        out[k] = t.numpy() if full else t[:, ::ROW_STRIDE * 4, :].numpy()
    for k in ("tokens", "scores", "P", "Pg", "Wa", "ba"):
        out["chk_" + k] = np.float64(synth.checksum(case[k]))
    out["chk_labels"] = np.float64(synth.checksum(case["labels"].float()))
    out["meta"] = np.array([shape.B, shape.N, shape.Din, shape.D, shape.K, shape.P, shape.Pg, shape.C, seed,
                            ROW_STRIDE], dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  logits[0,:3]={ref['logits'][0, :3].tolist()} "
          f"near_tie={len(out['near_tie'])}")


def main():
    if not ref_harness.reference_available():
        raise SystemExit("reference tree not available; fixtures can only be produced in the build container")
    only = set(sys.argv[1:])
    for name, key, b, seed, mode, fn in CASES:
        if only and name not in only:
            continue
        shape = synth.SHAPES[key]
        if b is not None:
            shape = shape.with_batch(b)
        pack(name, shape, seed, mode, fn)


if __name__ == "__main__":
    main()
