"""Probe of the two ways to obtain peer-mapped buffers on this box (torch symmetric memory, CUDA IPC)."""
import os
import sys
import traceback

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(local, p) for p in range(world) if p != local], flush=True)
try:
    import torch.distributed._symmetric_memory as sm
    t = sm.empty(1 << 20, dtype=torch.float32, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "symm_mem ok: ptrs", [hex(p) for p in h.buffer_ptrs], "signal", [hex(p) for p in h.signal_pad_ptrs],
          "pad", h.signal_pad_size, "multicast", h.has_multicast_support, hex(h.multicast_ptr), flush=True)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer read", peer[:2].tolist(), flush=True)
    h.barrier()
except Exception:
    traceback.print_exc()
try:
    buf = torch.full((1 << 20,), float(rank + 1), device=dev)
    hd = buf.untyped_storage()._share_cuda_()
    allh = [None] * world
    dist.all_gather_object(allh, (hd, buf.storage_offset(), buf.numel()))
    p = (rank + 1) % world
    hp, off, n = allh[p]
    (sdev, handle, ssize, soff, ref_handle, ref_off, ev_handle, ev_sync) = hp
    st = torch.UntypedStorage._new_shared_cuda(local, handle, ssize, soff, ref_handle, ref_off, ev_handle, ev_sync)
    peer = torch.empty(0, dtype=torch.float32, device=dev).set_(st, off, (n,))
    torch.cuda.synchronize()
    dist.barrier()
    print(rank, "ipc peer read", peer[:2].tolist(), hex(peer.data_ptr()), flush=True)
    dist.barrier()
except Exception:
    traceback.print_exc()
sys.stdout.flush()
os._exit(0)
