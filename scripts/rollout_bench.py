"""Attention-rollout -> CLS-row score (next #1): device time of pph_rollout_scores (CUDA events, inputs larger than L2)
against the HBM roofline, next to the CPU port of the reference's full-product rollout.  Synthetic attention maps."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import rollout_oracle as R
from protopformer_b200 import ops

dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
cases = [(11, 64, 3, 197), (11, 256, 3, 197), (11, 64, 6, 197), (24, 64, 4, 196)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in c.split(",")) for c in sys.argv[1].split(";")]
for L, B, H, T in cases:
    g = torch.Generator(device=dev).manual_seed(0)
    attn = [torch.softmax(2.0 * torch.randn(B, H, T, T, device=dev, generator=g), dim=-1) for _ in range(L)]
    nbytes = L * B * H * T * T * 4
    for _ in range(3):
        s = ops.rollout_scores(attn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        s = ops.rollout_scores(attn)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    row = dict(L=L, B=B, H=H, T=T, input_mb=round(nbytes / 1e6, 1), us=round(us, 1),
               gbs=round(nbytes / us / 1e3, 1), hbm_frac=round(nbytes / us / 1e3 / HBM, 3),
               images_per_s=round(B / us * 1e6))
    if B <= 64 and "--no-cpu" not in sys.argv:
        cb = min(B, 8)
        host = [a[:cb].cpu() for a in attn]
        torch.set_num_threads(os.cpu_count() or 1)
        R.rollout_full(host)
        t0 = time.perf_counter(); R.rollout_full(host); dt = time.perf_counter() - t0
        row.update(cpu_images_per_s=round(cb / dt, 1), cpu_cores=torch.get_num_threads())
        want = R.rollout_cls_row(host)
        row["max_rel_vs_oracle"] = float(((s[:cb].cpu() - want).abs() / want.abs().clamp_min(1e-12)).max())
    print(json.dumps(row), flush=True)
    del attn
