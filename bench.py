#!/usr/bin/env python
"""Benchmark of the prototype-head path (BASELINE.json metric: prototype-head images/sec fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|bf16] [--impl ours|reference]

A step = one training pass of the head over one per-GPU batch of BASELINE.json configs[1] (CUB / DeiT-Ti shape,
batch 64 per GPU): selection -> add-on -> similarity/pool -> last layers -> PPC loss -> cross-entropy -> backward
(token, prototype and add-on gradients) [-> gradient all-reduce when N > 1].  Synthetic inputs, random-init
parameters (oracle/synth.py).  Prints ONE JSON line (rank 0).

  value        device-resident throughput: inputs already in HBM, CUDA-graph replays, CUDA events, max over ranks.
               The step rotates over NBUF distinct input batches whose total size exceeds L2 (config.l2).
  e2e          same metric through the public API with HOST (pinned) inputs: H2D copy of every step's batch and a
               D2H read of the loss inside the timed region, double-buffered on a copy stream.
  roofline     the dominant kernel (tcgen05 similarity), timed alone with CUDA events on its launch stream.
  cpu_baseline the oracle's ATen-call-faithful port of the reference head (oracle/protohead_oracle.RefStyleHead)
               timed on this box's host cores on a bounded sample (rank 0, N=1 only).
  --impl reference : that CPU port as its own arm (the reference is pure Python over ATen; /root/reference does not
               exist on the GPU box, so the port stands in: cpu_baseline.kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "prototype_head_train_images_per_sec"
UNIT = "images/s"
WORKLOAD = "cub_b64"            # BASELINE.json configs[1]: deit_tiny CUB shape, training step, batch 64 per GPU


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML from a thread: the timed region is short)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.window = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, [k for k, b in names.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        t0, t1 = self.window or (0, float("inf"))
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        reasons = sorted({r for s in inside for r in s[2]})
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(inside)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference head's ATen call sequence on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_port_rate(shape, budget_s: float, sample_B: int, iters_min: int = 3):
    """images/s of RefStyleHead.train_step on a `sample_B`-image slice of the workload, within ~budget_s."""
    from oracle import protohead_oracle as O
    from protopformer_b200 import synth
    s = shape.with_batch(sample_B)
    case = synth.make_case(s, seed=1)
    head = O.RefStyleHead(case, s)
    head.train_step(case["tokens"], case["scores"], case["labels"])          # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < iters_min or (time.perf_counter() < t_end and len(times) < 50):
        t = time.perf_counter()
        head.train_step(case["tokens"], case["scores"], case["labels"])
        times.append(time.perf_counter() - t)
    return sample_B / statistics.median(times), len(times)


def run_reference_arm(args):
    from oracle import protohead_oracle as O
    from protopformer_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shape = synth.SHAPES[WORKLOAD]
    # bounded sample: a per-step slice of the 64-image batch sized so that (K + W) steps end within ~150 s
    probe, _ = cpu_port_rate(shape, 2.0, 8)
    per_step = 150.0 / max(1, args.steps + args.warmup)
    sample_B = max(1, min(shape.B, int(probe * per_step)))
    s = shape.with_batch(sample_B)
    case = synth.make_case(s, seed=1)
    head = O.RefStyleHead(case, s)
    for _ in range(args.warmup):
        head.train_step(case["tokens"], case["scores"], case["labels"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        head.train_step(case["tokens"], case["scores"], case["labels"])
    dt = time.perf_counter() - t0
    value = sample_B * args.steps / dt
    sample = f"{sample_B} of {shape.B} images per step ({WORKLOAD}), fp32, torch {torch.__version__} CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": shape.B, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def measure_next_rows(shape, params, dev, peaks):
    """Device time of the two HBM-bound rows next to the head, each against the measured HBM bandwidth:
    attention rollout -> CLS-row score (DeiT-Ti: 11 layers x 3 heads x 197^2 per image, read once) and the fused AdamW
    step over the head's parameter groups (28 B per element)."""
    from protopformer_b200 import ops
    from protopformer_b200.optim import FusedHeadAdamW
    out = {}
    with torch.no_grad():
        L, H, T, B = 11, 3, shape.N + 1, shape.B
        g = torch.Generator(device=dev).manual_seed(0)
        attn = [torch.softmax(2.0 * torch.randn(B, H, T, T, device=dev, generator=g), dim=-1) for _ in range(L)]
        for _ in range(3):
            ops.rollout_scores(attn)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(10):
            ops.rollout_scores(attn)
        r1.record()
        torch.cuda.synchronize()
        us = 1e3 * r0.elapsed_time(r1) / 10
        nbytes = float(L * B * H * T * T * 4)
        out["rollout"] = {"kernel": "rollout_prepare2_kernel + rollout_chain2_kernel", "bound": "hbm",
                          "achieved": nbytes / us / 1e3, "peak": peaks["hbm"], "unit": "GB/s",
                          "frac": nbytes / us / 1e3 / peaks["hbm"], "us_per_call": us,
                          "algorithmic_bytes_per_call": nbytes,
                          "workload": f"L={L} H={H} T={T} B={B} fp32 maps ({nbytes / 1e6:.0f} MB > L2), read once"}
        del attn
        ps = [params[k].detach().clone() for k in ("Wa", "ba", "P", "Pg")]
        gs = [torch.randn_like(p) * 0.01 for p in ps]
        opt = FusedHeadAdamW([{"params": ps[:2], "lr": 3e-3, "weight_decay": 1e-3},
                              {"params": ps[2:], "lr": 3e-3, "weight_decay": 0.05}], grads=gs)
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                opt.step()
        gr.replay()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            gr.replay()
        a1.record()
        torch.cuda.synchronize()
        us = 1e3 * a0.elapsed_time(a1) / 200
        nbytes = 28.0 * sum(p.numel() for p in ps)
        out["adamw"] = {"kernel": "adamw_kernel (4 tensors, one launch)", "bound": "hbm",
                        "achieved": nbytes / us / 1e3, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": nbytes / us / 1e3 / peaks["hbm"], "us_per_call": us,
                        "algorithmic_bytes_per_call": nbytes,
                        "note": "22.5 MB working set is L2-resident between replays: an upper bound on the HBM rate"}
    return out


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"],
                    help="similarity precision: fp32 = 3-term bf16 split (1e-4 parity), bf16 = single pass")
    ap.add_argument("--nbuf", type=int, default=16, help="distinct input batches the step rotates over")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph-allreduce", action="store_true",
                    help="N>1 (experimental, unverified): record the NCCL gradient all-reduce inside the CUDA graph "
                         "instead of issuing it after each replay")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    from protopformer_b200 import synth                       # synthetic input generator (bench infrastructure)
    from protopformer_b200 import _lib, ops
    from protopformer_b200.graph import GraphedHeadStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the prototype head has no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    shape = synth.SHAPES[WORKLOAD]
    case = synth.make_case(shape, seed=1)
    params = {k: case[k].to(dev) for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
    for k in ("Wa", "ba", "P", "Pg"):
        params[k].requires_grad_(True)
    cfg = ops.HeadConfig(K=shape.K, global_coe=shape.global_coe, mode=args.mode,
                         ppc_cov_thresh=shape.ppc_cov_thresh, ppc_mean_thresh=shape.ppc_mean_thresh)
    nbuf = args.nbuf
    step = GraphedHeadStep(params, cfg, B=shape.B, N=shape.N, C=shape.C, m=shape.m, n_slots=nbuf,
                           allreduce_in_graph=(world > 1 and args.graph_allreduce))
    # distinct synthetic batches per slot and per rank (weak scaling: fixed per-GPU batch)
    host = []
    for i in range(nbuf):
        c = synth.make_case(shape, seed=100 + rank * nbuf + i) if i < 4 else None
        if c is None:       # cheaper variants of the first four: permute images (keeps every tensor distinct)
            b = host[i % 4]
            perm = torch.randperm(shape.B, generator=torch.Generator().manual_seed(i))
            c = dict(tokens=b["tokens"][perm].contiguous(), scores=b["scores"][perm].contiguous(),
                     labels=b["labels"][perm].contiguous())
        host.append({k: c[k].pin_memory() for k in ("tokens", "scores", "labels")})
        step.load(i, host[i]["tokens"], host[i]["scores"], host[i]["labels"])
    torch.cuda.synchronize()
    step.capture()
    in_bytes = sum(host[0][k].numel() * host[0][k].element_size() for k in host[0])
    l2_note = f"inputs rotate over {nbuf} batches = {nbuf * in_bytes / 1e6:.0f} MB > 126 MB L2; parameters stay L2-resident"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident timing ------------------------------------------------------------------------------
    def one(i):
        step.run(i % nbuf)
        if world > 1:
            step.allreduce_grads()

    for i in range(args.warmup):
        one(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        one(i)
    ev1.record()
    barrier()
    t_w1 = time.perf_counter()
    sampler.window = (t_w0, t_w1)
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * shape.B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ---------------------------------------------------
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(1).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        for i in range(n):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[s])                 # slot s no longer read by step i-2
                b = host[i % nbuf]
                step.load(s, b["tokens"], b["scores"], b["labels"])
                ready[s].record(copy_stream)
            cur.wait_event(ready[s])
            step.run(s)
            if world > 1:
                step.allreduce_grads()
            loss_host.copy_(step.loss[s].reshape(1), non_blocking=True)
            done[s].record(cur)
        torch.cuda.synchronize()
        return float(loss_host.item())

    for s in range(2):
        done[s].record(torch.cuda.current_stream())
    e2e_loop(max(3, args.warmup // 10))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last_loss = e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * shape.B * args.steps / (ms_e2e * 1e-3)
    sampler.stop()

    # ---- roofline of the dominant kernel (tcgen05 similarity), timed alone on its stream ---------------------------
    peaks = _peaks()
    roof = None
    dominant = None
    if rank == 0:
        with torch.no_grad():
            tfs = []
            for i in range(nbuf):
                idx = ops.select_topk(step.scores[i], shape.K)
                tfs.append(ops.addon(step.tokens[i].detach(), idx, params["Wa"].detach(), params["ba"].detach(), True))
            pl = ops.prepare_prototypes(params["P"].detach(), True)
            pg = ops.prepare_prototypes(params["Pg"].detach(), True)
            for i in range(20):
                ops._similarity_raw(cfg, tfs[i % nbuf], pl, pg)
            torch.cuda.synchronize()
            reps = 200
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for i in range(reps):
                ops._similarity_raw(cfg, tfs[i % nbuf], pl, pg)
            k1.record()
            torch.cuda.synchronize()
            # launches are back to back on one stream (Python launch overhead < kernel time is NOT guaranteed for a
            # ~10 us kernel), so also time a CUDA-graph of the same launches and keep the smaller per-launch time
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(nbuf):
                    ops._similarity_raw(cfg, tfs[i], pl, pg)
            g.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(10):
                g.replay()
            g1.record()
            torch.cuda.synchronize()
            us_eager = 1e3 * k0.elapsed_time(k1) / reps
            us_graph = 1e3 * g0.elapsed_time(g1) / (10 * nbuf)
            us = min(us_eager, us_graph)
        flops = shape.B * (2.0 * shape.K * shape.D * shape.P + 2.0 * shape.D * shape.Pg)   # F_sim, SURVEY.md 8(d)
        achieved = flops / (us * 1e-6) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the committed ncu --set full
        # capture (profiles/r1_ncu_full_step_kernels.txt, cold caches); its operands total 7.1 MB, i.e. no re-reads
        traffic = 7.22e6 if (WORKLOAD == "cub_b64" and args.mode == "fp32") else None
        roof = {"kernel": "similarity_tc2_kernel (tcgen05 + TMA, resident prototype tile, mode %s)" % args.mode,
                "bound": "tensor", "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_burst"], "traffic": traffic, "us_per_launch": us,
                "algorithmic_flops_per_launch": flops,
                "peak_source": peaks["src"] + " bf16 burst (kernel timed alone)",
                "note": ("fp32 mode issues 3 bf16 MMA passes per algorithmic flop: tensor-pipe work is 3x the algorithmic "
                         "figure" if args.mode == "fp32" else "")}
        # the kernel with the largest share of the step (profiles/r1_launches_final_warm.txt): the argmin-routed sparse
        # backward, a byte/latency-bound gather.  Timed alone the same way; roofline = HBM with its compulsory bytes.
        try:
            with torch.no_grad():
                f = step.fused
                B_, K_, D_, P_, Pg_ = shape.B, shape.K, shape.D, shape.P, shape.Pg
                args_b = (f.g_l, f.g_g, f.argmin, f.Zs, f.Zc, params["P"].detach(), params["Pg"].detach(), B_, K_, D_, P_,
                          Pg_, f.ws, 2, None, None, f.dZs, f.dZc, torch.empty_like(params["P"]),
                          torch.empty_like(params["Pg"]))
                for _ in range(3):
                    _lib.call("pph_similarity_bwd", *args_b)
                torch.cuda.synchronize()
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb):
                    for _ in range(20):
                        _lib.call("pph_similarity_bwd", *args_b)
                gb.replay()
                torch.cuda.synchronize()
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                for _ in range(10):
                    gb.replay()
                b1.record()
                torch.cuda.synchronize()
                us_b = 1e3 * b0.elapsed_time(b1) / 200
            # compulsory bytes: Zs, Zc, P, Pg, g_l, g_g, argmin, bins read once; dZs, dZc, dP, dPg written once
            byts = 4.0 * (2 * B_ * K_ * D_ + 2 * B_ * D_ + 2 * (P_ + Pg_) * D_ + B_ * (P_ + Pg_) + 2 * B_ * P_)
            ach = byts / (us_b * 1e-6) / 1e9
            dominant = {"kernel": "sim_grads_kernel (argmin-routed sparse backward: dP, dPg, dZs, dZc)", "bound": "hbm",
                        "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                        "us_per_launch": us_b, "algorithmic_bytes_per_launch": byts,
                        "note": "gathers 2*B*P rows of D floats from L2-resident operands (197 MB of L2 traffic): "
                                "latency/L2-bound, far from the HBM roofline by construction"}
        except Exception as exc:      # the headline numbers must not depend on this auxiliary measurement
            dominant = {"error": str(exc)}

    # ---- rows either side of the head: rollout -> score (HBM bound), fused AdamW (HBM bound) --------
    next_rows = None
    if rank == 0:
        try:
            next_rows = measure_next_rows(shape, params, dev, peaks)
        except Exception as exc:      # auxiliary: the headline line must print regardless
            next_rows = {"error": str(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        rate, n = cpu_port_rate(shape, 12.0, shape.B)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} training steps of the full {shape.B}-image batch ({WORKLOAD}), fp32, torch CPU"}
        try:                              # per-core figure: same port on ONE thread, small sample
            torch.set_num_threads(1)
            rate1, n1 = cpu_port_rate(shape, 4.0, 8, iters_min=2)
            cpu["value_1_core"] = rate1
            cpu["sample_1_core"] = f"{n1} training steps of an 8-image slice, 1 thread"
        except Exception as exc:
            cpu["value_1_core"] = None
            cpu["sample_1_core"] = f"failed: {exc}"
        finally:
            torch.set_num_threads(cores)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split, fp32 accumulate (fp32-grade)" if args.mode == "fp32" else "bf16, fp32 accumulate",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": shape.B, "tokens": shape.K, "dim": shape.D,
                       "prototypes": shape.P, "global_prototypes": shape.Pg, "classes": shape.C, "mode": args.mode,
                       "step": "head fwd + PPC + CE + bwd (dtokens, dP, dPg, dWa, dba)"
                               + ((" + NCCL grad all-reduce (" + ("in graph" if args.graph_allreduce else "after each replay") + ")")
                                  if world > 1 else ""),
                       "l2": l2_note, "parallelism": f"dp{world}", "cuda_graph": True},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss},
            "gpu_launches": step.kernel_launches_per_step * args.steps,
            "gpu_launches_per_step": step.kernel_launches_per_step,
            "clocks": sampler.summary(),
            "roofline": roof,
            "largest_share_kernel": dominant if rank == 0 else None,
            "next_rows": next_rows,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
