"""Import the UNMODIFIED reference head (/root/reference/protopformer.py) on CPU  --  TEST INFRASTRUCTURE ONLY.

Used in the build container (where /root/reference exists) by tests/golden/make_golden.py to produce the
golden fixtures, and by tests/test_oracle_vs_reference.py (skipped when the reference tree is absent, e.g. on
the GPU box).  Nothing here is shipped or timed as product.

What has to be shimmed:
  * timm (pinned 0.5.4, README.md:59) and tkinter (`from turtle import forward`, tools/deit_models_attn.py:1)
    are not installed -> empty stub modules are registered before the import.
  * the head hard-codes `.cuda()` (protopformer.py:202,262,264,269,285) -> Tensor.cuda is patched to identity.
  * the backbone is replaced by a fake features module that returns preset tokens / CLS-attention scores; its
    class must be called MyVisionTransformer (protopformer.py:78-81).
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("PROTOPFORMER_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "protopformer.py"))


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Returns the reference's `protopformer` module (imported once)."""
    if "protopformer" in sys.modules and hasattr(sys.modules["protopformer"], "PPNet"):
        return sys.modules["protopformer"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if "timm" not in sys.modules:
        timm = _stub("timm")
        models = _stub("timm.models")
        timm.models = models
        models.create_model = lambda *a, **k: None
        vt = _stub("timm.models.vision_transformer")
        vt.VisionTransformer = type("VisionTransformer", (nn.Module,), {})
        vt._cfg = lambda **k: {}
        _stub("timm.models.registry").register_model = lambda f: f
        layers = _stub("timm.models.layers")
        for n in ("trunc_normal_", "PatchEmbed", "Mlp", "DropPath"):
            setattr(layers, n, object)
        _stub("timm.models.cait").Cait = type("Cait", (nn.Module,), {})
        helpers = _stub("timm.models.helpers")
        helpers.build_model_with_cfg = helpers.overlay_external_default_cfg = None
        data = _stub("timm.data")
        data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
        data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    if "turtle" not in sys.modules:
        _stub("turtle").forward = None
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF_ROOT)
    try:
        import protopformer  # noqa: the reference's own module
    finally:
        sys.path.remove(REF_ROOT)
    return protopformer


class MyVisionTransformer(nn.Module):
    """Fake backbone: hands the head preset `tokens` (B,1+N,Din) and `attn` (B,N)."""

    def __init__(self, din: int, n: int):
        super().__init__()
        self.fc = nn.Linear(din, din)          # the head reads .out_features of the last nn.Linear
        self.patch_embed = nn.Module()
        self.patch_embed.num_patches = n
        self.tokens = None
        self.attn = None

    def forward_feature_patch_embed_all(self, x):
        return self.tokens[:, :1], self.tokens[:, 1:]

    def forward_feature_mask_train_direct(self, cls_embed, x_embed, token_attn, reserve_layer_nums):
        return self.tokens, (self.attn, None)


def build_reference_head(case: dict, shape, fn: str = "log"):
    """Reference PPNet with parameters overwritten by the synthetic case (so RNG order inside the ctor is moot)."""
    ref = import_reference()
    s = shape
    feats = MyVisionTransformer(s.Din, s.N)
    net = ref.PPNet(feats, 224, [s.P, s.D, 1, 1], [14, 16, 16, 8.0], s.C,
                    reserve_layers=[11], reserve_token_nums=[s.K], use_global=True, use_ppc_loss=True,
                    ppc_cov_thresh=s.ppc_cov_thresh, ppc_mean_thresh=s.ppc_mean_thresh, global_coe=s.global_coe,
                    global_proto_per_class=s.Pg // s.C, prototype_activation_function=fn,
                    add_on_layers_type="regular")
    with torch.no_grad():
        net.prototype_vectors.copy_(case["P"].reshape(s.P, s.D, 1, 1))
        net.prototype_vectors_global.copy_(case["Pg"].reshape(s.Pg, s.D, 1, 1))
        net.add_on_layers[0].weight.copy_(case["Wa"].reshape(s.D, s.Din, 1, 1))
        net.add_on_layers[0].bias.copy_(case["ba"])
        net.last_layer.weight.copy_(case["Wl"])
        net.last_layer_global.weight.copy_(case["Wg"])
    return net, feats


def run_reference(case: dict, shape, fn: str = "log", cov_coe: float = 0.1, mean_coe: float = 0.5) -> dict:
    """Eval forward, push_forward, train forward + PPC + CE + backward through the reference's own code."""
    s = shape
    net, feats = build_reference_head(case, s, fn)
    dummy = torch.zeros(s.B, 3, 8, 8)
    feats.tokens, feats.attn = case["tokens"], case["scores"]
    out = {}
    net.eval()
    with torch.no_grad():
        logits, (attn, dist, lg, ll) = net(dummy)
        _, proto_acts = net.push_forward(dummy)
    out.update(logits=logits, logits_global=lg, logits_local=ll, dist_map=dist.flatten(2),
               act_map=proto_acts.flatten(2))
    net.train()
    tokens = case["tokens"].clone().requires_grad_(True)
    feats.tokens = tokens
    logits_t, aux = net(dummy)
    assert aux[0] is None and aux[4] == s.N
    cov_l, mean_l = net.get_PPC_loss(aux[2], aux[3], aux[4], case["labels"])
    ce = torch.nn.functional.cross_entropy(logits_t, case["labels"])
    loss = ce + cov_coe * cov_l + mean_coe * mean_l
    loss.backward()
    out.update(logits_train=logits_t.detach(), ce=ce.detach(), ppc_cov=cov_l.detach(), ppc_mean=mean_l.detach(),
               loss=loss.detach(), g_tokens=tokens.grad, g_P=net.prototype_vectors.grad.reshape(s.P, s.D),
               g_Pg=net.prototype_vectors_global.grad.reshape(s.Pg, s.D),
               g_Wa=net.add_on_layers[0].weight.grad.reshape(s.D, s.Din), g_ba=net.add_on_layers[0].bias.grad)
    # quantities the reference only holds implicitly
    am = out["act_map"]
    out["act_l"], out["argmax"] = am.max(dim=-1)
    out["dmin_l"] = out["dist_map"].min(dim=-1).values
    out["idx"] = torch.topk(case["scores"], k=s.K, dim=-1)[1].sort(dim=-1)[0]
    return out
