"""CPU oracle for ProtoPFormer's prototype head  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors, fp32 or float64) of the reference algorithm
in ``/root/reference/protopformer.py``.  It is the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it;
the product package ``protopformer_b200`` never does (tests/test_host_cpu.py::test_product_never_imports_the_oracle enforces that).

Parity pin: the reference ships no tests or golden vectors for this path.  The pin is the
reference's own code, imported unmodified in the build container by ``tests/golden/make_golden.py`` (through
``oracle/ref_harness.py``) to produce the fixtures under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those fixtures.  The arithmetic itself lives in PyTorch ATen
(conv2d/topk/sort/gather/max_pool2d/cdist; reference pins pytorch==1.8.1, README.md:56; this image has 2.11).

Reference map (file:line are in /root/reference/protopformer.py):
  select_tokens        : 157-162   topk -> sort -> gather
  addon                : 164-172, ctor 109-113  ('regular' add-on: 1x1 conv + sigmoid)
  l2_distances         : 201-218   relu(x2 - 2 x.p + p2)
  similarity           : 228-234   log((d+1)/(d+eps)) | -d
  pooled_activations   : 236-247   max over the token map
  logits               : 295-300 / 314-316
  ppc_loss             : 249-288   (batch_cov + get_PPC_loss)
  head_eval/head_train : 290-335
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPSILON = 1e-4  # protopformer.py:41


# ----------------------------------------------------------------------------------------------------------------
# a1: selection  (protopformer.py:157-162)
# ----------------------------------------------------------------------------------------------------------------
def select_tokens(scores: torch.Tensor, K: int) -> torch.Tensor:
    """scores (B,N) or (B,H,N) -> ascending index list (B,K) int64 of the K largest (head-mean) scores."""
    if scores.dim() == 3:
        scores = scores.mean(dim=1)
    top = torch.topk(scores, k=K, dim=-1).indices      # :157
    return top.sort(dim=-1).values                       # :158


def select_tokens_by_rank(scores: torch.Tensor, K: int) -> torch.Tensor:
    """Pure restatement without topk: n is selected  <=>  #{j : s_j > s_n} < K (tie-free inputs). SURVEY §8(d)(ii)."""
    if scores.dim() == 3:
        scores = scores.mean(dim=1)
    rank = (scores[:, None, :] > scores[:, :, None]).sum(dim=-1)           # (B,N): how many beat n
    sel = rank < K
    N = scores.shape[-1]
    ar = torch.arange(N).expand_as(scores)
    return ar[sel].reshape(scores.shape[0], K)


# ----------------------------------------------------------------------------------------------------------------
# a2: gather + add-on layer  (protopformer.py:159-172)
# ----------------------------------------------------------------------------------------------------------------
def addon(tokens: torch.Tensor, idx: torch.Tensor, Wa: torch.Tensor, ba: torch.Tensor):
    """tokens (B,1+N,Din), idx (B,K) -> Zs (B,K,D) selected-token features, Zc (B,D) CLS features.

    The reference permutes to NCHW and applies a 1x1 conv + sigmoid; a 1x1 conv is a per-token linear map.
    """
    B, _, Din = tokens.shape
    img = tokens[:, 1:]
    sel = torch.gather(img, 1, idx[:, :, None].expand(-1, -1, Din))          # :159-162
    Zs = torch.sigmoid(sel @ Wa.t() + ba)                                    # :172
    Zc = torch.sigmoid(tokens[:, 0] @ Wa.t() + ba)                           # :171
    return Zs, Zc


# ----------------------------------------------------------------------------------------------------------------
# a3-a5: distances, similarity, pooling  (protopformer.py:201-247)
# ----------------------------------------------------------------------------------------------------------------
def l2_distances(Z: torch.Tensor, P: torch.Tensor) -> torch.Tensor:
    """Z (B,K,D), P (Pn,D) -> d (B,Pn,K) = relu(|z|^2 - 2 z.p + |p|^2)  (:201-218, same association order)."""
    x2 = (Z * Z).sum(dim=-1)                      # (B,K)      :204-205
    p2 = (P * P).sum(dim=-1)                      # (Pn,)      :207-208
    xp = torch.einsum("bkd,pd->bpk", Z, P)        #            :213
    inter = -2.0 * xp + p2[None, :, None]         #            :214
    return F.relu(x2[:, None, :] + inter)         #            :216


def similarity(d: torch.Tensor, fn: str = "log", eps: float = EPSILON) -> torch.Tensor:
    if fn == "log":
        return torch.log((d + 1.0) / (d + eps))   # :230
    if fn == "linear":
        return -d                                 # :232
    raise ValueError(fn)


def pooled_activations(Z: torch.Tensor, P: torch.Tensor, fn: str = "log", eps: float = EPSILON, route=None):
    """-> act (B,Pn) max over tokens, argmax (B,Pn), dmin (B,Pn), full activation map (B,Pn,K), distances.

    `route` (B,Pn) int64, optional: pool at the given token instead of the arg-max.  Used by the gradient tests
    to make the oracle back-propagate through the token the kernel picked when the top-2 distance gap is below
    fp32 resolution (SURVEY.md §7 "near-tie policy"); with route == argmax it is the reference computation.
    """
    d = l2_distances(Z, P)
    a = similarity(d, fn, eps)
    if route is None:
        act, arg = a.max(dim=-1)                  # max_pool2d over the whole map, :242-244
        dmin = d.min(dim=-1).values
    else:
        arg = route
        act = torch.gather(a, 2, route[:, :, None]).squeeze(-1)
        dmin = torch.gather(d, 2, route[:, :, None]).squeeze(-1)
    return act, arg, dmin, a, d


# ----------------------------------------------------------------------------------------------------------------
# a6: last layers + combine  (protopformer.py:297-300)
# ----------------------------------------------------------------------------------------------------------------
def logits_from_activations(act_l, act_g, Wl, Wg, global_coe: float):
    lg = act_g @ Wg.t()
    ll = act_l @ Wl.t()
    return global_coe * lg + (1.0 - global_coe) * ll, lg, ll


def class_activation_maps(act_map: torch.Tensor, idx: torch.Tensor, labels: torch.Tensor, m: int, N: int):
    """eval_interpretability.py:195-225: gather the m prototypes of each image's label from the (B,P,K) activation map
    (:198-202) and scatter the K reserved-token activations to the grid of all N tokens (:218-223) -> (B,m,N)."""
    B, _, K = act_map.shape
    rows = labels[:, None] * m + torch.arange(m)[None, :]
    sel = torch.gather(act_map, 1, rows[:, :, None].expand(-1, -1, K))
    out = torch.zeros(B, m, N, dtype=act_map.dtype)
    out.scatter_(2, idx[:, None, :].expand(-1, m, -1), sel)
    return out


# ----------------------------------------------------------------------------------------------------------------
# a7: PPC loss  (protopformer.py:249-288)
# ----------------------------------------------------------------------------------------------------------------
def ppc_loss(act_map: torch.Tensor, idx: torch.Tensor, labels: torch.Tensor, m: int, N: int,
             cov_thresh: float, mean_thresh: float):
    """act_map (B,P,K) activations, idx (B,K) ascending selected tokens, labels (B,) -> (L_cov, L_mean).

    Restated per SURVEY §8(d)(iii): weights live on the K selected grid cells (zero elsewhere, :276), the
    grid position of token n is (n // side, n % side) (:262), N = side*side is the ORIGINAL token count.
    """
    B, _, K = act_map.shape
    side = int(round(math.sqrt(N)))
    rows = (labels[:, None] * m + torch.arange(m)[None, :])                           # (B,m)   :268-269
    w = torch.gather(act_map, 1, rows[:, :, None].expand(-1, -1, K))                   # (B,m,K) :271
    pos = torch.stack([idx // side, idx % side], dim=-1).to(w.dtype)                   # (B,K,2) :262
    S = w.sum(dim=-1, keepdim=True)                                                    # (B,m,1)
    wn = w / S * N                                                                     # :251
    mean = (pos[:, None, :, :] * wn[..., None]).sum(dim=2) / N                         # (B,m,2) :252 (mean over N cells)
    # cells outside the selection carry weight 0 -> contribute nothing to the weighted sums
    diff = pos[:, None, :, :] - mean[:, :, None, :]                                    # (B,m,K,2) :253
    var = (wn[..., None] * diff * diff).sum(dim=2) / (N - 1)                           # (B,m,2) diag of :254-256
    cov_l = F.relu((var[..., 0] + var[..., 1]) / 2.0 - cov_thresh).mean()              # :280-281
    dist = torch.cdist(mean, mean)                                                     # (B,m,m) :284
    mask = 1.0 - torch.eye(m, dtype=w.dtype)
    mean_l = F.relu((mean_thresh - dist) * mask).mean()                                # :286
    return cov_l, mean_l


# ----------------------------------------------------------------------------------------------------------------
# a6/a9: whole-head entry points (what PPNet.forward / push_forward / get_PPC_loss compute after the backbone)
# ----------------------------------------------------------------------------------------------------------------
def head_forward(case: dict, K: int, global_coe: float, fn: str = "log", eps: float = EPSILON, route=None) -> dict:
    """case: tokens, scores (or scores_h), P, Pg, Wa, ba, Wl, Wg -> every intermediate the tests compare."""
    scores = case.get("scores_h", case["scores"])
    idx = select_tokens(scores, K)
    Zs, Zc = addon(case["tokens"], idx, case["Wa"], case["ba"])
    act_l, arg_l, dmin_l, amap, dmap = pooled_activations(Zs, case["P"], fn, eps, route)
    act_g, _, dmin_g, _, _ = pooled_activations(Zc[:, None, :], case["Pg"], fn, eps)
    logits, lg, ll = logits_from_activations(act_l, act_g, case["Wl"], case["Wg"], global_coe)
    return dict(idx=idx, Zs=Zs, Zc=Zc, act_l=act_l, argmax=arg_l, dmin_l=dmin_l, act_map=amap, dist_map=dmap,
                act_g=act_g, dmin_g=dmin_g, logits=logits, logits_global=lg, logits_local=ll)


def head_train_step(case: dict, shape, ppc_cov_coe: float = 0.1, ppc_mean_coe: float = 0.5,
                    fn: str = "log", eps: float = EPSILON, dtype=torch.float32, route=None) -> dict:
    """Forward + CE + PPC + backward (loss = CE + cov_coe*L_cov + mean_coe*L_mean, scripts/train_cub.sh:43-44;
    engine_proto.py:51-64).  Gradients by autograd on this restatement.  Returns outputs and grads."""
    c = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in case.items()}
    for k in ("tokens", "P", "Pg", "Wa", "ba"):
        c[k].requires_grad_(True)
    out = head_forward(c, shape.K, shape.global_coe, fn, eps, route)
    ce = F.cross_entropy(out["logits"], c["labels"])
    cov_l, mean_l = ppc_loss(out["act_map"], out["idx"], c["labels"], shape.m, shape.N,
                             shape.ppc_cov_thresh, shape.ppc_mean_thresh)
    loss = ce + ppc_cov_coe * cov_l + ppc_mean_coe * mean_l
    loss.backward()
    res = {k: v.detach() for k, v in out.items()}
    res.update(ce=ce.detach(), ppc_cov=cov_l.detach(), ppc_mean=mean_l.detach(), loss=loss.detach(),
               g_tokens=c["tokens"].grad, g_P=c["P"].grad, g_Pg=c["Pg"].grad, g_Wa=c["Wa"].grad, g_ba=c["ba"].grad)
    return res


# ----------------------------------------------------------------------------------------------------------------
# ATen-call-faithful variant used ONLY as the timed CPU baseline ("port"): it issues the same library calls the
# reference issues (conv2d against an all-ones filter bank, conv2d against the prototypes, elementwise chain,
# max_pool2d, second topk+sort and scatter_ in the PPC loss, bmm outer products, cdist) so its cost on the host
# cores is the reference's cost.  protopformer.py:141-172, 201-218, 228-247, 249-288, 290-335.
# ----------------------------------------------------------------------------------------------------------------
class RefStyleHead(torch.nn.Module):
    def __init__(self, case: dict, shape, fn: str = "log"):
        super().__init__()
        s = shape
        self.s, self.fn = s, fn
        self.prototype_vectors = torch.nn.Parameter(case["P"].clone().reshape(s.P, s.D, 1, 1))
        self.prototype_vectors_global = torch.nn.Parameter(case["Pg"].clone().reshape(s.Pg, s.D, 1, 1))
        self.conv = torch.nn.Conv2d(s.Din, s.D, 1)
        with torch.no_grad():
            self.conv.weight.copy_(case["Wa"].reshape(s.D, s.Din, 1, 1))
            self.conv.bias.copy_(case["ba"])
        self.last_layer = torch.nn.Linear(s.P, s.C, bias=False)
        self.last_layer_global = torch.nn.Linear(s.Pg, s.C, bias=False)
        with torch.no_grad():
            self.last_layer.weight.copy_(case["Wl"])
            self.last_layer_global.weight.copy_(case["Wg"])
        self.last_layer.weight.requires_grad = False
        self.last_layer_global.weight.requires_grad = False

    def _features(self, tokens, scores):
        s = self.s
        idx = torch.topk(scores, k=s.K, dim=-1)[1].sort(dim=-1)[0]
        wide = idx[:, :, None].repeat(1, 1, s.Din)                       # the materialised int64 index (:159)
        cls, img = tokens[:, :1], tokens[:, 1:]
        img = torch.gather(img, 1, wide)
        side = int(round(math.sqrt(s.K)))
        cls = cls.permute(0, 2, 1).reshape(-1, s.Din, 1, 1)
        img = img.permute(0, 2, 1).reshape(-1, s.Din, side, side)
        return torch.sigmoid(self.conv(cls)), torch.sigmoid(self.conv(img))

    def _acts(self, x, protos):
        ones = torch.ones(protos.shape)                                  # re-allocated per call (:202)
        x2 = F.conv2d(x * x, ones)
        p2 = (protos ** 2).sum(dim=(1, 2, 3)).view(-1, 1, 1)
        d = F.relu(x2 + (-2 * F.conv2d(x, protos) + p2))
        a = torch.log((d + 1) / (d + EPSILON)) if self.fn == "log" else -d
        full = a
        if a.shape[-1] > 1:
            a = F.max_pool2d(a, kernel_size=(a.shape[-1], a.shape[-1]))
        return a.reshape(x.shape[0], protos.shape[0]), d, full

    def forward(self, tokens, scores):
        s = self.s
        zc, zs = self._features(tokens, scores)
        ag, _, _ = self._acts(zc, self.prototype_vectors_global)
        al, d, full = self._acts(zs, self.prototype_vectors)
        lg, ll = self.last_layer_global(ag), self.last_layer(al)
        return s.global_coe * lg + (1 - s.global_coe) * ll, d, full, lg, ll

    def ppc(self, full, scores, labels):
        s = self.s
        side = int(round(math.sqrt(s.N)))
        B, m = full.shape[0], s.m
        grid = torch.FloatTensor([[x, y] for x in range(side) for y in range(side)])
        grid = grid[None].repeat(B * m, 1, 1)
        wts = torch.zeros(B, m, s.N)
        flat = full.flatten(start_dim=2)
        rows = (labels * m).unsqueeze(-1).repeat(1, m) + torch.arange(m)
        sel = torch.gather(flat, 1, rows[:, :, None].repeat(1, 1, s.K))
        idx = torch.topk(scores, k=s.K, dim=-1)[1].sort(dim=-1)[0][:, None, :].repeat(1, m, 1)
        wts.scatter_(2, idx, sel)
        wts = wts.reshape(B * m, -1)
        wts = wts / wts.sum(dim=-1, keepdim=True) * s.N
        mean = (grid * wts[:, :, None]).mean(dim=1).unsqueeze(1)
        diffs = (grid - mean).reshape(B * m * s.N, 2)
        prods = torch.bmm(diffs.unsqueeze(2), diffs.unsqueeze(1)).reshape(B * m, s.N, 2, 2)
        cov = (prods * wts[:, :, None, None]).sum(dim=1) / (s.N - 1)
        cov_l = F.relu((cov[:, 0, 0] + cov[:, 1, 1]) / 2 - s.ppc_cov_thresh).mean()
        mean = mean.reshape(B, m, 2)
        mean_l = F.relu((s.ppc_mean_thresh - torch.cdist(mean, mean)) * (1. - torch.eye(m))).mean()
        return cov_l, mean_l

    def train_step(self, tokens, scores, labels, cov_coe=0.1, mean_coe=0.5):
        for p in self.parameters():
            p.grad = None
        tokens = tokens.detach().requires_grad_(True)       # the backbone receives d(loss)/d(tokens)
        logits, _, full, _, _ = self.forward(tokens, scores)
        loss = F.cross_entropy(logits, labels)
        cov_l, mean_l = self.ppc(full, scores, labels)
        loss = loss + cov_coe * cov_l + mean_coe * mean_l
        loss.backward()
        return loss.detach()
