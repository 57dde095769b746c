"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Differentiable torch (CPU) restatement of the reference's LPD-Net training
path AS WRITTEN, used as the gradient oracle at sizes without a committed golden and as the CPU baseline of the C3
(training step) bench line.  It is never imported by the package.

Reference lines followed (relative to the reference root):
  knn                    util/lpdnet_model.py:317-326   matmul + topk, as written
  get_graph_feature      util/lpdnet_model.py:331-363   edge = cat(neighbour, centre), materialised [B, 2C, N, k]
  LPDNet.forward         util/lpdnet_model.py:211-268   (t3d = tfea = False, as PointNetVlad constructs it by default)
  NetVLADLoupe.forward   util/PointNetVlad.py:45-83 ;  GatingContext.forward util/PointNetVlad.py:103-115
  quadruplet_loss        loss/pointnetvlad_loss.py:49-97 (+ best_pos_distance :6-12)
  run_model              train_pointnetvlad.py:202-217 (tuple layout)
Pinned by tests/test_oracle_vs_golden.py against the reference's own outputs / autograd gradients (tests/golden/c3_train_step_*).
Weights: a state_dict with the reference's key names; BatchNorm uses batch statistics when train=True.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def knn(x, k):
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return pd.topk(k=k, dim=-1)[1]


def graph_feature(x, k, idx):
    B, C, N = x.shape
    base = torch.arange(B).view(-1, 1, 1) * N
    flat = (idx + base).reshape(-1)
    xt = x.transpose(2, 1).contiguous()
    feat = xt.reshape(B * N, C)[flat].view(B, N, k, C)
    centre = xt.view(B, N, 1, C).expand(-1, -1, k, -1)
    return torch.cat((feat, centre), dim=3).permute(0, 3, 1, 2)


def _bn(x, sd, p, train, stats=None):
    rm, rv = sd[p + "running_mean"], sd[p + "running_var"]
    if train:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[p + "weight"], sd[p + "bias"], training=train, momentum=0.1, eps=1e-5)
    if stats is not None and train:
        stats[p + "running_mean"], stats[p + "running_var"] = rm, rv
    return y


def lpdnet_forward(sd, x, k=20, train=True, prefix="emb_nn.", slope=0.01, stats=None):
    """x [B, 1, N, 3] -> [B, emb, N, 1]"""
    act = lambda t: F.leaky_relu(t, slope)
    x = x.squeeze(1).transpose(2, 1)                                         # [B, 3, N]
    xinit = x
    p = prefix
    x = act(_bn(F.conv1d(x, sd[p + "conv1_lpd.weight"]), sd, p + "bn1_lpd.", train, stats))
    x = act(_bn(F.conv1d(x, sd[p + "conv2_lpd.weight"]), sd, p + "bn2_lpd.", train, stats))
    idx = knn(x.detach(), k)
    e = graph_feature(x, k, idx)
    y = act(_bn(F.conv2d(e, sd[p + "convDG1.0.weight"]), sd, p + "convDG1.1.", train, stats))
    x1 = y.max(dim=-1, keepdim=True)[0]
    y = act(_bn(F.conv2d(y, sd[p + "convDG2.0.weight"]), sd, p + "convDG2.1.", train, stats))
    x2 = y.max(dim=-1, keepdim=True)[0]
    idx = knn(xinit.detach(), k)
    e = graph_feature(x2.squeeze(-1), k, idx)
    y = act(_bn(F.conv2d(e, sd[p + "convSN1.0.weight"]), sd, p + "convSN1.1.", train, stats))
    x3 = y.max(dim=-1, keepdim=True)[0]
    cat = torch.cat((x1, x2, x3), dim=1).squeeze(-1)
    x = act(_bn(F.conv1d(cat, sd[p + "conv3_lpd.weight"]), sd, p + "bn3_lpd.", train, stats))
    return x.unsqueeze(-1)


def netvlad_forward(sd, x, train=True, prefix="net_vlad.", stats=None):
    """x [B, D, N, 1] -> [B, out]"""
    p = prefix
    B, D, N = x.shape[:3]
    x = x.transpose(1, 3).contiguous().view(B, N, D)
    K = sd[p + "cluster_weights"].shape[1]
    a = torch.matmul(x, sd[p + "cluster_weights"]).view(-1, K)
    a = _bn(a, sd, p + "bn1.", train, stats).view(B, N, K)
    a = F.softmax(a, dim=-1)
    a_sum = a.sum(-2, keepdim=True) * sd[p + "cluster_weights2"]
    v = torch.matmul(a.transpose(2, 1), x).transpose(2, 1) - a_sum
    v = F.normalize(v, dim=1, p=2).reshape(B, K * D)
    v = F.normalize(v, dim=1, p=2)
    h = _bn(torch.matmul(v, sd[p + "hidden1_weights"]), sd, p + "bn2.", train, stats)
    g = _bn(torch.matmul(h, sd[p + "context_gating.gating_weights"]), sd, p + "context_gating.bn1.", train, stats)
    return h * torch.sigmoid(g)


def pointnetvlad_forward(sd, x, k=20, train=True, stats=None):
    return netvlad_forward(sd, lpdnet_forward(sd, x, k, train, stats=stats), train, stats=stats)


def quadruplet_loss(q, pos, neg, other, m1, m2):
    """hot configuration: use_min=True, lazy=True, ignore_zero_loss=False"""
    P, Nn = pos.shape[1], neg.shape[1]
    dpos = ((pos - q.expand(-1, P, -1)) ** 2).sum(2).min(1)[0].view(-1, 1)
    l1 = (m1 + dpos.expand(-1, Nn) - ((neg - q.expand(-1, Nn, -1)) ** 2).sum(2)).clamp(min=0.0).max(1)[0].mean()
    l2 = (m2 + dpos.expand(-1, Nn) - ((neg - other.expand(-1, Nn, -1)) ** 2).sum(2)).clamp(min=0.0).max(1)[0].mean()
    return l1 + l2


def train_step(sd, x, Bq, P=2, Nn=18, m1=0.5, m2=0.2, k=20):
    """one forward + loss + backward on tuples in run_model order; returns (out, loss, {key: grad}, running stats)"""
    leaf = {key: (v.clone().requires_grad_(True) if v.is_floating_point() and not key.endswith(("running_mean", "running_var")) else v)
            for key, v in sd.items()}
    stats = {}
    out = pointnetvlad_forward(leaf, x, k, True, stats)
    o = out.view(Bq, -1, out.shape[1])
    q, pos, neg, other = torch.split(o, [1, P, Nn, 1], dim=1)
    loss = quadruplet_loss(q, pos, neg, other, m1, m2)
    loss.backward()
    grads = {key: v.grad for key, v in leaf.items() if getattr(v, "grad", None) is not None}
    return out.detach(), loss.detach(), grads, stats
