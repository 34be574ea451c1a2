"""Drop-in module level: protopformer_b200.PPNet behind the reference's forward / get_PPC_loss / push_forward
signatures (engine_proto.py:49-59, 162-179), driven by the same fake backbone the golden fixtures were made with."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.util import load_golden, norm_rel, rel_close

pytestmark = pytest.mark.gpu


class MyVisionTransformer(nn.Module):
    """Fake backbone handing the head preset tokens / CLS-attention (the class NAME matters, protopformer.py:78-81)."""

    def __init__(self, din, n):
        super().__init__()
        self.fc = nn.Linear(din, din)
        self.patch_embed = nn.Module()
        self.patch_embed.num_patches = n
        self.tokens = self.attn = None

    def forward_feature_patch_embed_all(self, x):
        return self.tokens[:, :1], self.tokens[:, 1:]

    def forward_feature_mask_train_direct(self, cls_embed, x_embed, token_attn, reserve_layer_nums):
        return self.tokens, (self.attn, None)


def _build(shape, case, fn, precision):
    from protopformer_b200 import PPNet
    feats = MyVisionTransformer(shape.Din, shape.N)
    net = PPNet(feats, 224, [shape.P, shape.D, 1, 1], [14, 16, 16, 8.0], shape.C, reserve_layers=[11],
                reserve_token_nums=[shape.K], use_global=True, use_ppc_loss=True,
                ppc_cov_thresh=shape.ppc_cov_thresh, ppc_mean_thresh=shape.ppc_mean_thresh,
                global_coe=shape.global_coe, global_proto_per_class=shape.Pg // shape.C,
                prototype_activation_function=fn, add_on_layers_type='regular', precision=precision)
    with torch.no_grad():
        net.prototype_vectors.copy_(case["P"].reshape(shape.P, shape.D, 1, 1))
        net.prototype_vectors_global.copy_(case["Pg"].reshape(shape.Pg, shape.D, 1, 1))
        net.add_on_layers[0].weight.copy_(case["Wa"].reshape(shape.D, shape.Din, 1, 1))
        net.add_on_layers[0].bias.copy_(case["ba"])
        net.last_layer.weight.copy_(case["Wl"])
        net.last_layer_global.weight.copy_(case["Wg"])
    return net.cuda(), feats


@pytest.mark.parametrize("name,precision", [("cub_b8_s1", "fp32"), ("small_s1", "fp32"), ("tiny_s2_linear", "fp32"),
                                            ("cars_b4_s1", "fp32_fma")])
def test_ppnet_dropin_eval_train_push(name, precision):
    shape, case, g, fn = load_golden(name)
    net, feats = _build(shape, case, fn, precision)
    dummy = torch.zeros(shape.B, 3, 8, 8, device="cuda")
    feats.tokens, feats.attn = case["tokens"].cuda(), case["scores"].cuda()
    # ---- eval: (logits, (cls_token_attn, distances, logits_global, logits_local)), protopformer.py:301
    net.eval()
    with torch.no_grad():
        logits, aux = net(dummy)
    assert len(aux) == 4 and aux[0] is feats.attn
    assert rel_close(logits.cpu(), g["logits"], 1e-4)
    assert rel_close(aux[2].cpu(), g["logits_global"], 1e-4) and rel_close(aux[3].cpu(), g["logits_local"], 1e-4)
    side = int(round(shape.K ** 0.5))
    assert tuple(aux[1].shape) == (shape.B, shape.P, side, side)
    stride = 1 if shape.name in ("tiny", "small") else int(g["meta"][9]) * 4
    assert rel_close(aux[1].flatten(2).cpu()[:, ::stride], g["dist_map"], 1e-4)
    # ---- push_forward: (cls_token_attn, proto_acts (B,P,h,w)), protopformer.py:344
    attn, acts = net.push_forward(dummy)
    assert attn is feats.attn and tuple(acts.shape) == (shape.B, shape.P, side, side)
    assert rel_close(acts.flatten(2).cpu()[:, ::stride], g["act_map"], 1e-4)
    # ---- train: (logits, (None, zeros(1), total_proto_act, rollout, 196)) + get_PPC_loss, engine_proto.py:49-64
    net.train()
    tokens = case["tokens"].cuda().requires_grad_(True)
    feats.tokens = tokens
    logits_t, aux = net(dummy)
    assert aux[0] is None and aux[1].shape == (1,) and aux[4] == shape.N and not aux[3].requires_grad
    labels = case["labels"].cuda()
    cov, mean = net.get_PPC_loss(aux[2], aux[3], aux[4], labels)
    assert cov.dim() == 0 and mean.dim() == 0
    loss = F.cross_entropy(logits_t, labels) + 0.1 * cov + 0.5 * mean
    loss.backward()
    assert rel_close(loss.cpu(), g["loss"], 1e-4)
    assert rel_close(cov.cpu(), g["ppc_cov"], 1e-4) and rel_close(mean.cpu(), g["ppc_mean"], 1e-4)
    assert net.last_layer.weight.grad is None and net.last_layer_global.weight.grad is None and net.ones.grad is None
    if len(g["near_tie"]) == 0:
        full = shape.name in ("tiny", "small")
        st = 1 if full else int(g["meta"][9])
        gp = net.prototype_vectors.grad.reshape(shape.P, shape.D).cpu()
        assert norm_rel(gp[::st], g["g_P"]) < 2e-4
        assert norm_rel(net.add_on_layers[0].weight.grad.reshape(shape.D, shape.Din).cpu()[::st], g["g_Wa"]) < 2e-4
        assert norm_rel(tokens.grad.reshape(-1, shape.Din).cpu()[::st], g["g_tokens"]) < 2e-4


@pytest.mark.parametrize("name", ["cub_b8_s1", "small_s1", "cars_b4_s1"])
def test_get_ppc_loss_accepts_a_dense_map_tensor(name):
    """protopformer.py:259-271 takes any (B,P,h,w) tensor: the dense-map kernels must give the reference fixture's losses
    and the oracle's gradient with respect to that tensor."""
    from oracle import protohead_oracle as O
    shape, case, g, fn = load_golden(name)
    net, feats = _build(shape, case, fn, "fp32")
    net.train()
    labels = case["labels"].cuda()
    ref = O.head_forward(case, shape.K, shape.global_coe, fn)
    side = int(round(shape.K ** 0.5))
    dense_cpu = ref["act_map"].reshape(shape.B, shape.P, side, side).clone().requires_grad_(True)
    cov_o, mean_o = O.ppc_loss(dense_cpu.flatten(2), ref["idx"], case["labels"], shape.m, shape.N, shape.ppc_cov_thresh,
                               shape.ppc_mean_thresh)
    (0.3 * cov_o + 0.7 * mean_o).backward()
    dense = ref["act_map"].reshape(shape.B, shape.P, side, side).cuda().requires_grad_(True)
    cov, mean = net.get_PPC_loss(dense, case["scores"].cuda(), shape.N, labels)
    (0.3 * cov + 0.7 * mean).backward()
    assert rel_close(cov.cpu(), g["ppc_cov"], 1e-4) and rel_close(mean.cpu(), g["ppc_mean"], 1e-4)
    assert rel_close(cov.cpu(), cov_o.detach(), 1e-5) and rel_close(mean.cpu(), mean_o.detach(), 1e-5)
    assert norm_rel(dense.grad.cpu(), dense_cpu.grad) < 1e-4
    rows = (case["labels"][:, None] * shape.m + torch.arange(shape.m)[None]).long()
    mask = torch.ones(shape.B, shape.P, dtype=torch.bool)
    mask[torch.arange(shape.B)[:, None], rows] = False
    assert (dense.grad.cpu()[mask] == 0).all()                       # only the label-class rows receive gradient
    with pytest.raises(TypeError):
        net.get_PPC_loss([1, 2, 3], case["scores"].cuda(), shape.N, labels)


def test_ppnet_under_autocast_and_grad_scaler():
    """The reference runs the head under torch.cuda.amp.autocast with a loss scaler (engine_proto.py:48,76-77):
    the custom ops keep their own fp32 cast policy and scale linearly."""
    shape, case, g, fn = load_golden("cub_b8_s1")
    net, feats = _build(shape, case, fn, "fp32")
    net.train()
    dummy = torch.zeros(shape.B, 3, 8, 8, device="cuda")
    feats.tokens, feats.attn = case["tokens"].cuda().half(), case["scores"].cuda()
    labels = case["labels"].cuda()
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    with torch.autocast("cuda", dtype=torch.float16):
        logits, aux = net(dummy)
        cov, mean = net.get_PPC_loss(aux[2], aux[3], aux[4], labels)
        loss = F.cross_entropy(logits, labels) + 0.1 * cov + 0.5 * mean
    assert logits.dtype == torch.float32
    scaler.scale(loss).backward()
    gp = net.prototype_vectors.grad
    assert torch.isfinite(gp).all() and gp.abs().max() > 0
    # fp16 tokens cost ~1e-3 on the logits; the point is that the path runs and stays finite
    assert rel_close(logits.float().cpu(), g["logits_train"], 5e-3)


def test_state_dict_keys_match_reference_checkpoint_layout():
    shape, case, g, fn = load_golden("tiny_s1")
    net, _ = _build(shape, case, fn, "fp32")
    keys = {k for k in net.state_dict().keys() if not k.startswith("features.")}
    assert keys == {"prototype_vectors", "prototype_vectors_global", "ones", "add_on_layers.0.weight",
                    "add_on_layers.0.bias", "last_layer.weight", "last_layer_global.weight"}
