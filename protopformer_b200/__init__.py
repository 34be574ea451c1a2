"""protopformer_b200 -- ProtoPFormer's prototype head as sm_100a CUDA kernels behind the reference's PPNet API.

    from protopformer_b200 import PPNet, construct_PPNet      # drop-in for protopformer.py
    from protopformer_b200 import ops                          # functional head: ops.head_forward / ops.ppc_loss
    from protopformer_b200 import GraphedHeadStep              # the CUDA-graph training / inference step (graph.py)

The compute lives in lib/libprotohead_b200.so (C ABI: include/protohead.h), built by `python -m protopformer_b200.build`.
"""
from . import ops  # noqa: F401
from .head import PPNet, ProtoMap, construct_PPNet, base_architecture_to_features  # noqa: F401
from .ops import HeadConfig, head_forward, ppc_loss, ppc_loss_dense, select_topk  # noqa: F401
from .graph import GraphedHeadStep  # noqa: F401
from .dist import FlatGradReducer, PeerGradReducer, shard_batch  # noqa: F401

__all__ = ["PPNet", "ProtoMap", "construct_PPNet", "base_architecture_to_features", "HeadConfig", "head_forward",
           "ppc_loss", "ppc_loss_dense", "select_topk", "ops", "GraphedHeadStep", "FlatGradReducer", "PeerGradReducer",
           "shard_batch"]
