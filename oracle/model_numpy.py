"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference forward pass, AS WRITTEN
(channel-major tensors, materialised [B, 2C, N, k] edge tensors, BatchNorm not folded), so that it checks
the fused / decomposed kernels against the reference's own formulation.

Reference lines followed (relative to the reference root):
  knn                     util/lpdnet_model.py:317-326   (canonical order, see knn_canonical.c)
  get_graph_feature       util/lpdnet_model.py:331-363   edge = cat(neighbour, centre)
  get_graph_feature_Origin util/lpdnet_model.py:116-145  edge = cat(centre, neighbour - centre) / gather only
  LPDNet.forward          util/lpdnet_model.py:211-268
  LPDNetOrign.forward     util/lpdnet_model.py:64-114
  TranformNet.forward     util/lpdnet_model.py:295-313
  STN3d.forward           util/PointNetVlad.py:149-179
  PointNetfeat.forward    util/PointNetVlad.py:203-241
  NetVLADLoupe.forward    util/PointNetVlad.py:45-83
  GatingContext.forward   util/PointNetVlad.py:103-115
  PointNetVlad.forward    util/PointNetVlad.py:261-270

Weights come in as a state_dict of numpy float32 arrays with the reference's key names.
`train=True` uses batch statistics in every BatchNorm (biased variance, eps 1e-5) like module.train().
"""
from __future__ import annotations

import numpy as np

from . import knn_canonical

EPS = 1e-5
f32 = np.float32


def _np(sd):
    return {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in sd.items()}


def batchnorm(x, sd, prefix, train):
    """BatchNorm over axis 1 (channels) of x [B, C, ...]."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    if train:
        axes = tuple(a for a in range(x.ndim) if a != 1)
        mean = x.astype(np.float64).mean(axis=axes)
        var = x.astype(np.float64).var(axis=axes)  # biased, as used for normalisation
        mean, var = mean.astype(f32), var.astype(f32)
    else:
        mean, var = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    y = (x - mean.reshape(shape)) / np.sqrt(var.reshape(shape) + f32(EPS))
    return (y * sd[prefix + "weight"].reshape(shape) + sd[prefix + "bias"].reshape(shape)).astype(f32)


def conv1x1(x, w, bias=None):
    """x [B, Cin, ...] ; w [Cout, Cin, (1(,1))] -> [B, Cout, ...]"""
    w2 = w.reshape(w.shape[0], -1)
    B, Cin = x.shape[:2]
    y = np.matmul(w2[None], x.reshape(B, Cin, -1))
    if bias is not None:
        y = y + bias.reshape(1, -1, 1)
    return y.reshape((B, w2.shape[0]) + x.shape[2:]).astype(f32)


def leaky(x, slope=0.01):
    return np.where(x > 0, x, x * f32(slope)).astype(f32)


def relu(x):
    return np.maximum(x, 0).astype(f32)


KNN_IMPL = "canonical"   # "canonical": bit-exact tie-by-index target (parity) | "blas": as written (timing baseline)


def knn_blas(x, k):
    """knn() exactly as written in the reference (:317-326): SGEMM inner product, two subtractions, top-k —
    BLAS accumulation order and unspecified tie order, like torch.  Used for the CPU timing baseline only."""
    inner = -2 * np.matmul(x.transpose(0, 2, 1), x)
    xx = np.sum(x ** 2, axis=1, keepdims=True)
    pd = -xx - inner
    pd = pd - xx.transpose(0, 2, 1)
    part = np.argpartition(-pd, k - 1, axis=-1)[..., :k]
    order = np.argsort(-np.take_along_axis(pd, part, -1), axis=-1, kind="stable")
    return np.take_along_axis(part, order, -1)


def knn(x, k):
    """x [B, C, N] -> idx [B, N, k] (canonical tie order unless KNN_IMPL == "blas")."""
    if KNN_IMPL == "blas":
        return knn_blas(x, k)
    return knn_canonical(np.ascontiguousarray(x.transpose(0, 2, 1)), k)


def _gather(x, idx):
    """x [B, C, N], idx [B, N, k] -> neighbour features [B, N, k, C]"""
    B = x.shape[0]
    xt = x.transpose(0, 2, 1)  # [B, N, C]
    return xt[np.arange(B)[:, None, None], idx]


def get_graph_feature(x, k=20, idx=None):
    if idx is None:
        idx = knn(x, k)
    feat = _gather(x, idx)                                     # [B, N, k, C]
    centre = np.broadcast_to(x.transpose(0, 2, 1)[:, :, None, :], feat.shape)
    return np.concatenate((feat, centre), axis=3).transpose(0, 3, 1, 2).astype(f32)   # [B, 2C, N, k]


def get_graph_feature_origin(x, k=20, idx=None, cat=True):
    if idx is None:
        idx = knn(x, k)
    feat = _gather(x, idx)
    if cat:
        centre = np.broadcast_to(x.transpose(0, 2, 1)[:, :, None, :], feat.shape)
        feat = np.concatenate((centre, feat - centre), axis=3)
    return feat.transpose(0, 3, 1, 2).astype(f32)


def tranform_net(sd, p, x, train, k):
    """TranformNet: x [B, k, N] -> [B, k, k]"""
    x = relu(batchnorm(conv1x1(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"]), sd, p + "bn1.", train))
    x = relu(batchnorm(conv1x1(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"]), sd, p + "bn2.", train))
    x = relu(batchnorm(conv1x1(x, sd[p + "conv3.weight"], sd[p + "conv3.bias"]), sd, p + "bn3.", train))
    x = x.max(axis=2)                                          # [B, 1024]
    x = relu(batchnorm(x @ sd[p + "fc1.weight"].T + sd[p + "fc1.bias"], sd, p + "bn4.", train))
    x = relu(batchnorm(x @ sd[p + "fc2.weight"].T + sd[p + "fc2.bias"], sd, p + "bn5.", train))
    x = x @ sd[p + "fc3.weight"].T + sd[p + "fc3.bias"]
    x = x + np.eye(k, dtype=f32).reshape(1, k * k)
    return x.reshape(-1, k, k).astype(f32)


def lpdnet_forward(sd, x, train=False, k=20, prefix="emb_nn.", per_cloud_eval=True):
    """LPDNet.forward: x [B, 1, N, 3] -> [B, emb, N, 1].  In eval mode clouds are independent and are
    processed one at a time to bound memory."""
    sd = _np(sd)
    if not train and per_cloud_eval and x.shape[0] > 1:
        return np.concatenate([lpdnet_forward(sd, x[b:b + 1], False, k, prefix) for b in range(x.shape[0])], axis=0)
    p = prefix
    x = np.asarray(x, dtype=f32)[:, 0].transpose(0, 2, 1)      # [B, 3, N]
    x_init = x
    if p + "t_net3d.conv1.weight" in sd:
        trans = tranform_net(sd, p + "t_net3d.", x, train, 3)
        x = np.matmul(x.transpose(0, 2, 1), trans).transpose(0, 2, 1)
    x = leaky(batchnorm(conv1x1(x, sd[p + "conv1_lpd.weight"]), sd, p + "bn1_lpd.", train))
    x = leaky(batchnorm(conv1x1(x, sd[p + "conv2_lpd.weight"]), sd, p + "bn2_lpd.", train))
    if p + "t_net_fea.conv1.weight" in sd:
        tf = tranform_net(sd, p + "t_net_fea.", x, train, 64)
        x = np.matmul(x.transpose(0, 2, 1), tf).transpose(0, 2, 1)
    e = get_graph_feature(x, k)                                # [B, 128, N, k]
    e = leaky(batchnorm(conv1x1(e, sd[p + "convDG1.0.weight"]), sd, p + "convDG1.1.", train))
    x1 = e.max(axis=-1, keepdims=True)
    e = leaky(batchnorm(conv1x1(e, sd[p + "convDG2.0.weight"]), sd, p + "convDG2.1.", train))
    x2 = e.max(axis=-1, keepdims=True)
    idx = knn(x_init, k)
    e = get_graph_feature(x2[..., 0], k, idx)                  # [B, 256, N, k]
    e = leaky(batchnorm(conv1x1(e, sd[p + "convSN1.0.weight"]), sd, p + "convSN1.1.", train))
    x3 = e.max(axis=-1, keepdims=True)
    x = np.concatenate((x1, x2, x3), axis=1)[..., 0]           # [B, 512, N]
    x = leaky(batchnorm(conv1x1(x, sd[p + "conv3_lpd.weight"]), sd, p + "bn3_lpd.", train))
    return x[..., None]


def lpdnetorigin_forward(sd, x, train=False, k=20, prefix="emb_nn.", per_cloud_eval=True):
    sd = _np(sd)
    if not train and per_cloud_eval and x.shape[0] > 1:
        return np.concatenate([lpdnetorigin_forward(sd, x[b:b + 1], False, k, prefix) for b in range(x.shape[0])], axis=0)
    p = prefix

    def seq(x, name):
        return leaky(batchnorm(conv1x1(x, sd[p + name + ".0.weight"]), sd, p + name + ".1.", train))

    x = np.asarray(x, dtype=f32)[:, 0].transpose(0, 2, 1)
    x_init = x
    if p + "t_net3d.conv1.weight" in sd:
        trans = tranform_net(sd, p + "t_net3d.", x, train, 3)
        x = np.matmul(x.transpose(0, 2, 1), trans).transpose(0, 2, 1)
    x = seq(seq(x, "conv1_lpd"), "conv2_lpd")
    if p + "t_net_fea.conv1.weight" in sd:
        tf = tranform_net(sd, p + "t_net_fea.", x, train, 64)
        x = np.matmul(x.transpose(0, 2, 1), tf).transpose(0, 2, 1)
    e = get_graph_feature_origin(x, k)
    e = seq(seq(e, "convDG1"), "convDG2")
    x = e.max(axis=-1, keepdims=True)
    idx = knn(x_init, k)
    e = get_graph_feature_origin(x[..., 0], k, idx, cat=False)
    e = seq(seq(e, "convSN1"), "convSN2")
    x = e.max(axis=-1)
    x = seq(seq(seq(x, "conv3_lpd"), "conv4_lpd"), "conv5_lpd")
    return x[..., None]


def stn3d(sd, p, x, k, num_points):
    """STN3d with use_bn=False (the only way PointNetfeat builds it): x [B, C, N, W] -> [B, k, k]"""
    B = x.shape[0]
    w1 = sd[p + "conv1.weight"]                                 # [64, C, 1, W]
    y = np.einsum("bcnw,ocw->bon", x, w1[:, :, 0, :]) + sd[p + "conv1.bias"].reshape(1, -1, 1)
    y = relu(y.astype(f32))
    y = relu(conv1x1(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"]))
    y = relu(conv1x1(y, sd[p + "conv3.weight"], sd[p + "conv3.bias"]))
    assert y.shape[2] == num_points
    y = y.max(axis=2)
    y = relu(y @ sd[p + "fc1.weight"].T + sd[p + "fc1.bias"])
    y = relu(y @ sd[p + "fc2.weight"].T + sd[p + "fc2.bias"])
    y = y @ sd[p + "fc3.weight"].T + sd[p + "fc3.bias"]
    y = y + np.eye(k, dtype=f32).reshape(1, k * k)
    return y.reshape(B, k, k).astype(f32)


def pointnetfeat_forward(sd, x, train=False, prefix="point_net.", feature_transform=False):
    """PointNetfeat.forward with max_pool=False: x [B, 1, N, 3] -> [B, emb, N, 1]"""
    sd = _np(sd)
    p = prefix
    x = np.asarray(x, dtype=f32)
    B, _, N, _ = x.shape
    trans = stn3d(sd, p + "stn.", x, 3, N)
    x = np.matmul(x[:, 0], trans)                               # [B, N, 3]
    x = x.reshape(B, 1, N, 3)
    w1 = sd[p + "conv1.weight"]
    y = np.einsum("bcnw,ocw->bon", x, w1[:, :, 0, :]) + sd[p + "conv1.bias"].reshape(1, -1, 1)
    y = relu(batchnorm(y.astype(f32), sd, p + "bn1.", train))
    y = relu(batchnorm(conv1x1(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"]), sd, p + "bn2.", train))
    if feature_transform:
        ft = stn3d(sd, p + "feature_trans.", y[..., None], 64, N)
        y = np.matmul(y.transpose(0, 2, 1), ft).transpose(0, 2, 1)
    y = relu(batchnorm(conv1x1(y, sd[p + "conv3.weight"], sd[p + "conv3.bias"]), sd, p + "bn3.", train))
    y = relu(batchnorm(conv1x1(y, sd[p + "conv4.weight"], sd[p + "conv4.bias"]), sd, p + "bn4.", train))
    y = batchnorm(conv1x1(y, sd[p + "conv5.weight"], sd[p + "conv5.bias"]), sd, p + "bn5.", train)
    return y[..., None]


def netvlad_forward(sd, x, train=False, prefix="net_vlad."):
    """NetVLADLoupe.forward (gating=True, add_batch_norm=True): x [B, D, N, 1] -> [B, out]"""
    sd = _np(sd)
    p = prefix
    x = np.ascontiguousarray(np.asarray(x, dtype=f32)[..., 0].transpose(0, 2, 1))   # [B, N, D]
    B, N, D = x.shape
    wc = sd[p + "cluster_weights"]
    K = wc.shape[1]
    act = (x.reshape(-1, D) @ wc).astype(f32)                   # [B*N, K]
    act = batchnorm(act, sd, p + "bn1.", train)
    act = act - act.max(axis=1, keepdims=True)
    act = np.exp(act)
    act = (act / act.sum(axis=1, keepdims=True)).astype(f32).reshape(B, N, K)
    a = act.sum(axis=1, keepdims=True) * sd[p + "cluster_weights2"]              # [B, D, K]
    vlad = np.matmul(act.transpose(0, 2, 1), x).transpose(0, 2, 1) - a           # [B, D, K]
    vlad = vlad / np.maximum(np.linalg.norm(vlad, axis=1, keepdims=True), 1e-12)
    vlad = vlad.reshape(B, D * K)
    vlad = vlad / np.maximum(np.linalg.norm(vlad, axis=1, keepdims=True), 1e-12)
    h = (vlad.astype(f32) @ sd[p + "hidden1_weights"]).astype(f32)
    h = batchnorm(h, sd, p + "bn2.", train)
    g = (h @ sd[p + "context_gating.gating_weights"]).astype(f32)
    g = batchnorm(g, sd, p + "context_gating.bn1.", train)
    g = 1.0 / (1.0 + np.exp(-g))
    return (h * g).astype(f32)


def pointnetvlad_forward(sd, x, featnet="lpdnet", train=False, k=20, feature_transform=False):
    """PointNetVlad.forward: x [B, 1, N, 3] -> [B, output_dim]"""
    sd = _np(sd)
    if featnet == "lpdnet":
        f = lpdnet_forward(sd, x, train, k)
    elif featnet == "lpdnetorigin":
        f = lpdnetorigin_forward(sd, x, train, k)
    elif featnet == "pointnet":
        f = pointnetfeat_forward(sd, x, train, feature_transform=feature_transform)
    else:
        raise ValueError("featnet error")
    return netvlad_forward(sd, f, train)
