import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protopformer_b200 import synth
from protopformer_b200 import ops, _lib
dev = torch.device("cuda:0")
s = synth.SHAPES["cub_b64"]
case = {k: v.to(dev) for k, v in synth.make_case(s, seed=1).items()}
lib = _lib.load()
lib.pph_debug_read.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
idx = ops.select_topk(case["scores"], s.K)
def dump(tag):
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    lib.pph_debug_read(buf)
    t = list(buf); t0 = t[0]
    print(tag, {i: (t[i] - t0) for i in range(32) if t[i] >= t0 and t[i] - t0 < 10**8})
with torch.no_grad():
    for it in range(3):
        tf = ops.addon(case["tokens"], idx, case["Wa"], case["ba"], True)
        dump(f"addon_fwd[{it}] ns since CTA start:")
    B, N, Din, D, K = s.B, s.N, s.Din, s.D, s.K
    dZs = torch.randn_like(tf.Zs) * 1e-3; dZc = torch.randn_like(tf.Zc) * 1e-3
    dWa = torch.empty_like(case["Wa"]); dba = torch.empty(D, device=dev); dtok = torch.empty_like(case["tokens"])
    ws = ops.addon_bwd_workspace(B, N, Din, D, K, dev)
    for it in range(2):
        _lib.call("pph_addon_bwd", case["tokens"], idx, case["Wa"], tf.Zs, tf.Zc, dZs, dZc, B, N, Din, D, K, ws, 3, dWa, dba, dtok)
        dump(f"addon_bwd (last = dX gemm)[{it}]:")
