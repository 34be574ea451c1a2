"""N ranks, one per GPU: the step with the gradient all-reduce recorded inside the CUDA graph (overlapped with the
add-on backward) must leave, on every rank, the average of the per-rank gradients of the same step run without the
exchange.  Launch:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protopformer_b200 import ops, synth  # noqa: E402
from protopformer_b200.graph import GraphedHeadStep  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    shape = synth.SHAPES["cub_b64"]
    case = synth.make_case(shape, seed=1)
    batch = synth.make_case(shape, seed=50 + rank)
    cfg = ops.HeadConfig(K=shape.K, global_coe=shape.global_coe, mode="fp32", ppc_cov_thresh=shape.ppc_cov_thresh,
                         ppc_mean_thresh=shape.ppc_mean_thresh)

    def make(in_graph):
        params = {k: case[k].to(dev) for k in ("Wa", "ba", "P", "Pg", "Wl", "Wg")}
        for k in ("Wa", "ba", "P", "Pg"):
            params[k].requires_grad_(True)
        st = GraphedHeadStep(params, cfg, B=shape.B, N=shape.N, C=shape.C, m=shape.m, n_slots=2,
                             allreduce_in_graph=in_graph)
        for s in range(2):
            st.load(s, batch["tokens"], batch["scores"], batch["labels"])
        torch.cuda.synchronize()
        st.capture()
        return st

    local_step = make(False)
    local_step.run(0)
    torch.cuda.synchronize()
    want = local_step.reducer.flat.clone()
    dist.all_reduce(want, op=dist.ReduceOp.AVG)
    graphed = make(True)
    for i in range(50):                       # repeated replays: no hang, no drift
        graphed.run(i % 2)
    torch.cuda.synchronize()
    got = graphed.reducer.flat
    err = float((got - want).abs().max() / want.abs().max())
    own = float((got - local_step.reducer.flat).abs().max() / want.abs().max())
    print(f"rank {rank}/{world}: in-graph all-reduce vs averaged local gradients: {err:.3e} (vs own gradients {own:.3e})",
          flush=True)
    assert err < 1e-6, err
    assert own > 1e-3, "ranks hold different batches: the averaged gradient must differ from the local one"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
