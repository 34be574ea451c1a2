"""Drop-in replacement of the reference's ``PPNet`` (protopformer.py:12-401) for the prototype-head path.

Same constructor arguments, attribute names (``features``, ``add_on_layers``, ``prototype_vectors``,
``prototype_vectors_global``, ``ones``, ``last_layer``, ``last_layer_global`` -- tools/create_optimizer.py:31-39 and
the checkpoint keys of main.py:393-407 depend on them), ``forward`` / ``get_PPC_loss`` / ``push_forward``
signatures and aux-tuple layouts (engine_proto.py:49-59, 162-179).  Everything after the backbone returns runs in
the sm_100a kernels of libprotohead_b200.so; there is no PyTorch or CPU fallback for that part.

Differences that are deliberate:
  * only ``use_global=True`` and ``add_on_layers_type='regular'`` are supported (the only configuration the
    reference scripts use and the only one that runs without crashing in the reference);
  * the (B,P,h,w) maps in the aux tuples (train aux[2] ``total_proto_act``, eval aux[1] ``distances``) are lazy
    ``ProtoMap`` handles: ``get_PPC_loss`` consumes the handle directly, and ``handle.materialize()`` /
    ``handle.flatten()`` / indexing produce the reference tensor for the visualisation scripts;
  * one extra ctor keyword ``precision`` ('fp32' = 3-term bf16 split on tcgen05, 'bf16', 'fp32_fma').
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops


class ProtoMap:
    """Lazy stand-in for a (B,P,h,w) prototype map (activations or distances)."""

    def __init__(self, cfg, tf, P, Pg, p2l, kind: str):
        self.cfg, self.tf, self.P, self.Pg, self.p2l, self.kind = cfg, tf, P, Pg, p2l, kind
        B, K, _ = tf.Zs.shape
        side = int(round(math.sqrt(K)))
        self.shape = torch.Size((B, P.shape[0], side, side))
        self._dense = None

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            dist, act = ops.materialize_maps(self.cfg, self.tf, self.P, self.Pg)
            self._dense = (act if self.kind == "act" else dist).reshape(self.shape)
        return self._dense

    # the handful of tensor methods the reference's consumers use on the map
    def flatten(self, *a, **k):
        return self.materialize().flatten(*a, **k)

    def __getitem__(self, item):
        return self.materialize()[item]

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def detach(self):
        return self.materialize()

    def cpu(self):
        return self.materialize().cpu()

    def dim(self):
        return 4

    @property
    def device(self):
        return self.tf.Zs.device


class PPNet(nn.Module):
    def __init__(self, features, img_size, prototype_shape, proto_layer_rf_info, num_classes,
                 reserve_layers=[], reserve_token_nums=[], use_global=False, use_ppc_loss=False,
                 ppc_cov_thresh=2., ppc_mean_thresh=2, global_coe=0.3, global_proto_per_class=10,
                 init_weights=True, prototype_activation_function='log', add_on_layers_type='bottleneck',
                 precision='fp32'):
        super().__init__()
        if not use_global:
            raise NotImplementedError("use_global=False crashes in the reference (protopformer.py:148-155); "
                                      "only use_global=True is supported")
        if add_on_layers_type == 'bottleneck':
            raise NotImplementedError("add_on_layers_type='bottleneck' is never selected by the reference scripts; "
                                      "only 'regular' (1x1 conv + sigmoid) is supported")
        if prototype_activation_function not in ('log', 'linear'):
            raise NotImplementedError("prototype_activation_function must be 'log' or 'linear'")
        self.img_size = img_size
        self.prototype_shape = prototype_shape
        self.num_prototypes = prototype_shape[0]
        self.num_classes = num_classes
        self.reserve_layers = reserve_layers
        self.reserve_token_nums = reserve_token_nums
        self.use_global = use_global
        self.use_ppc_loss = use_ppc_loss
        self.ppc_cov_thresh = ppc_cov_thresh
        self.ppc_mean_thresh = ppc_mean_thresh
        self.global_coe = global_coe
        self.global_proto_per_class = global_proto_per_class
        self.epsilon = 1e-4                                                   # protopformer.py:41
        self.reserve_layer_nums = list(zip(self.reserve_layers, self.reserve_token_nums))
        self.num_prototypes_global = self.num_classes * self.global_proto_per_class
        self.prototype_shape_global = [self.num_prototypes_global] + list(self.prototype_shape[1:])
        self.prototype_activation_function = prototype_activation_function
        self.precision = precision

        assert self.num_prototypes % self.num_classes == 0                    # protopformer.py:57
        self.num_prototypes_per_class = self.num_prototypes // self.num_classes
        ident = torch.zeros(self.num_prototypes, self.num_classes)
        ident[torch.arange(self.num_prototypes), torch.arange(self.num_prototypes) // self.num_prototypes_per_class] = 1
        self.prototype_class_identity = ident
        identg = torch.zeros(self.num_prototypes_global, self.num_classes)
        identg[torch.arange(self.num_prototypes_global),
               torch.arange(self.num_prototypes_global) // self.global_proto_per_class] = 1
        self.prototype_class_identity_global = identg

        self.proto_layer_rf_info = proto_layer_rf_info
        self.features = features
        name = str(self.features).upper()
        if not (name.startswith('MYVISION') or name.startswith('MYCAIT')):
            raise Exception('other base base_architecture NOT implemented')   # protopformer.py:86
        in_channels = [i for i in features.modules() if isinstance(i, nn.Linear)][-1].out_features
        self.num_patches = self.features.patch_embed.num_patches

        self.add_on_layers = nn.Sequential(
            nn.Conv2d(in_channels=in_channels, out_channels=self.prototype_shape[1], kernel_size=1),
            nn.Sigmoid())
        self.prototype_vectors = nn.Parameter(torch.rand(self.prototype_shape), requires_grad=True)
        self.prototype_vectors_global = nn.Parameter(torch.rand(self.prototype_shape_global), requires_grad=True)
        self.ones = nn.Parameter(torch.ones(self.prototype_shape), requires_grad=False)   # checkpoint key only
        self.last_layer = nn.Linear(self.num_prototypes, self.num_classes, bias=False)
        self.last_layer_global = nn.Linear(self.num_prototypes_global, self.num_classes, bias=False)
        self.last_layer.weight.requires_grad = False
        self.last_layer_global.weight.requires_grad = False
        self.all_attn_mask = None
        self.teacher_model = None
        self.scale = self.prototype_shape[1] ** -0.5
        if init_weights:
            self._initialize_weights()

    # ------------------------------------------------------------------------------------------------------------
    def _cfg(self) -> ops.HeadConfig:
        K = self.reserve_layer_nums[-1][1]
        D = self.prototype_shape[1]
        mode = self.precision
        if mode != 'fp32_fma' and not ops.tc_supported(D, K):
            mode = 'fp32_fma'          # shapes outside the tcgen05 kernel's build range run on the FP32-FMA kernel
        return ops.HeadConfig(K=K, global_coe=float(self.global_coe), act_fn=self.prototype_activation_function,
                              eps=float(self.epsilon), mode=mode, ppc_cov_thresh=float(self.ppc_cov_thresh),
                              ppc_mean_thresh=float(self.ppc_mean_thresh))

    def _backbone(self, x):
        """protopformer.py:149,155 -- the two backbone calls; returns tokens (B,1+N,Din), cls_token_attn (B,N)."""
        cls_embed, x_embed = self.features.forward_feature_patch_embed_all(x)
        tokens, (cls_token_attn, _) = self.features.forward_feature_mask_train_direct(
            cls_embed, x_embed, None, self.reserve_layer_nums)
        return tokens, cls_token_attn

    def head(self, tokens, cls_token_attn) -> ops.HeadOutput:
        """Everything between the backbone's return and the logits (protopformer.py:156-172, 311-316)."""
        conv = self.add_on_layers[0]
        return ops.head_forward(self._cfg(), tokens, cls_token_attn, conv.weight, conv.bias,
                                self.prototype_vectors, self.prototype_vectors_global,
                                self.last_layer.weight, self.last_layer_global.weight)

    def forward(self, x):
        tokens, cls_token_attn = self._backbone(x)
        out = self.head(tokens, cls_token_attn)
        cfg = self._cfg()
        if not self.training:                                                  # protopformer.py:292-301
            distances = ProtoMap(cfg, out.tf, self.prototype_vectors, self.prototype_vectors_global, out.p2l, "dist")
            return out.logits, (cls_token_attn, distances, out.logits_global, out.logits_local)
        cls_attn_rollout = cls_token_attn.detach()                             # protopformer.py:306
        total_proto_act = ProtoMap(cfg, out.tf, self.prototype_vectors, self.prototype_vectors_global, out.p2l, "act")
        attn_loss = torch.zeros(1, device=out.logits.device)
        original_fea_len = int(cls_attn_rollout.shape[-1])
        return out.logits, (None, attn_loss, total_proto_act, cls_attn_rollout, original_fea_len)

    def get_PPC_loss(self, total_proto_act, cls_attn_rollout, original_fea_len, label):
        """protopformer.py:259-288.  `total_proto_act` is normally the ProtoMap returned by forward() in training mode (the
        selected-token list it carries is the one the reference would recompute from `cls_attn_rollout`, :273-274, and the
        label-class slice is recomputed from the token features instead of being gathered from a (B,P,h,w) tensor).  A
        dense (B,P,h,w) tensor is accepted as in the reference: the loss and its gradient w.r.t. that tensor come from
        the dense-map kernels."""
        if not isinstance(total_proto_act, ProtoMap):
            if not torch.is_tensor(total_proto_act) or total_proto_act.dim() != 4:
                raise TypeError("get_PPC_loss expects the ProtoMap returned by forward() or a (B,P,h,w) tensor")
            return ops.ppc_loss_dense(self._cfg(), total_proto_act, cls_attn_rollout, label,
                                      self.num_prototypes_per_class, int(original_fea_len))
        pm = total_proto_act
        return ops.ppc_loss(pm.cfg, pm.tf, self.prototype_vectors, pm.p2l, label,
                            self.num_prototypes_per_class, int(original_fea_len))

    def push_forward(self, x):
        """protopformer.py:337-344 -> (cls_token_attn (B,N), proto_acts (B,P,h,w)) with the map materialised."""
        tokens, cls_token_attn = self._backbone(x)
        cfg = self._cfg()
        idx32 = ops.select_topk(cls_token_attn, cfg.K)
        conv = self.add_on_layers[0]
        with torch.no_grad():
            tf = ops.addon(tokens, idx32, conv.weight, conv.bias, False)
            pm = ProtoMap(cfg, tf, self.prototype_vectors, self.prototype_vectors_global, None, "act")
            return cls_token_attn, pm.materialize()

    def push_forward_class_maps(self, x, targets):
        """(cls_token_attn (B,N), maps (B, m, side, side)): the activation maps of each image's label-class prototypes
        on the grid of ALL tokens (zeros on pruned ones) -- what eval_interpretability.py:195-225 derives from
        ``push_forward`` by a gather and a scatter -- without ever forming the (B,P,h,w) map."""
        tokens, cls_token_attn = self._backbone(x)
        cfg = self._cfg()
        idx32 = ops.select_topk(cls_token_attn, cfg.K)
        conv = self.add_on_layers[0]
        with torch.no_grad():
            tf = ops.addon(tokens, idx32, conv.weight, conv.bias, False)
            maps = ops.class_activation_maps(cfg, tf, self.prototype_vectors, targets, self.num_prototypes_per_class,
                                             cls_token_attn.shape[-1])
        return cls_token_attn, maps

    # ------------------------------------------------------------------------------------------------------------
    def __repr__(self):
        return ('PPNet(\n\tfeatures: {},\n\timg_size: {},\n\tprototype_shape: {},\n\tproto_layer_rf_info: {},\n'
                '\tnum_classes: {},\n\tepsilon: {}\n)').format(self.features, self.img_size, self.prototype_shape,
                                                               self.proto_layer_rf_info, self.num_classes, self.epsilon)

    def set_last_layer_incorrect_connection(self, incorrect_strength):
        """protopformer.py:367-386: +1 on the prototype's own class, `incorrect_strength` elsewhere."""
        pos = torch.t(self.prototype_class_identity)
        self.last_layer.weight.data.copy_(1 * pos + incorrect_strength * (1 - pos))
        posg = torch.t(self.prototype_class_identity_global)
        self.last_layer_global.weight.data.copy_(1 * posg + incorrect_strength * (1 - posg))

    def _initialize_weights(self):
        for m in self.add_on_layers.modules():                                 # protopformer.py:388-395
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.set_last_layer_incorrect_connection(incorrect_strength=-0.5)


def construct_PPNet(base_architecture, pretrained=True, img_size=224, prototype_shape=(2000, 512, 1, 1),
                    num_classes=200, reserve_layers=[], reserve_token_nums=[], use_global=False, use_ppc_loss=False,
                    ppc_cov_thresh=1., ppc_mean_thresh=2., global_coe=0.5, global_proto_per_class=10,
                    prototype_activation_function='log', add_on_layers_type='bottleneck', features=None,
                    precision='fp32'):
    """protopformer.py:455-487.  The backbone is out of this path's scope: pass it
    as `features` (any module exposing the two backbone methods), or register a factory for `base_architecture` in
    `base_architecture_to_features`."""
    if features is None:
        if base_architecture not in base_architecture_to_features:
            raise KeyError(f"no backbone factory registered for {base_architecture!r}; pass features=...")
        features = base_architecture_to_features[base_architecture](pretrained=pretrained)
    proto_layer_rf_info = [14, 16, 16, 8.0]
    return PPNet(features=features, img_size=img_size, prototype_shape=prototype_shape,
                 proto_layer_rf_info=proto_layer_rf_info, num_classes=num_classes, reserve_layers=reserve_layers,
                 reserve_token_nums=reserve_token_nums, use_global=use_global, use_ppc_loss=use_ppc_loss,
                 ppc_cov_thresh=ppc_cov_thresh, ppc_mean_thresh=ppc_mean_thresh, global_coe=global_coe,
                 global_proto_per_class=global_proto_per_class, init_weights=True,
                 prototype_activation_function=prototype_activation_function,
                 add_on_layers_type=add_on_layers_type, precision=precision)


base_architecture_to_features = {}
