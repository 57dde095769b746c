import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
from lpdnet_b200 import ops, train
from lpdnet_b200.util.lpdnet_model import TranformNet
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ops.set_precision("fp32")
for seed, nd in ((64, False),):
  torch.manual_seed(seed); k = 64
  if True:
    B, N = 6, 200
    net = TranformNet(k).cuda().train()
    with torch.no_grad():
        for p in net.parameters(): p.add_(torch.randn_like(p) * 0.2)
    rows = torch.randn(B * N, k, device="cuda")
    Wt = torch.randn(B, k, k, device="cuda")
    # torch reference (CUDA autograd)
    x = rows.view(B, N, k).transpose(1, 2)
    h = F.relu(F.batch_norm(F.conv1d(x, net.conv1.weight, net.conv1.bias), None, None, net.bn1.weight, net.bn1.bias, True))
    h = F.relu(F.batch_norm(F.conv1d(h, net.conv2.weight, net.conv2.bias), None, None, net.bn2.weight, net.bn2.bias, True))
    h = F.relu(F.batch_norm(F.conv1d(h, net.conv3.weight, net.conv3.bias), None, None, net.bn3.weight, net.bn3.bias, True))
    g = h.max(2)[0]
    g = F.relu(F.batch_norm(F.linear(g, net.fc1.weight, net.fc1.bias), None, None, net.bn4.weight, net.bn4.bias, True))
    g = F.relu(F.batch_norm(F.linear(g, net.fc2.weight, net.fc2.bias), None, None, net.bn5.weight, net.bn5.bias, True))
    T = (F.linear(g, net.fc3.weight, net.fc3.bias) + torch.eye(k, device="cuda").view(1, -1)).view(B, k, k)
    (T * Wt).sum().backward()
    ref = {n: p.grad.clone() for n, p in net.named_parameters()}
    import copy
    n64 = copy.deepcopy(net).double()
    for p in n64.parameters(): p.grad = None
    x = rows.double().view(B, N, k).transpose(1, 2)
    h = F.relu(F.batch_norm(F.conv1d(x, n64.conv1.weight, n64.conv1.bias), None, None, n64.bn1.weight, n64.bn1.bias, True))
    h = F.relu(F.batch_norm(F.conv1d(h, n64.conv2.weight, n64.conv2.bias), None, None, n64.bn2.weight, n64.bn2.bias, True))
    h = F.relu(F.batch_norm(F.conv1d(h, n64.conv3.weight, n64.conv3.bias), None, None, n64.bn3.weight, n64.bn3.bias, True))
    g = h.max(2)[0]
    g = F.relu(F.batch_norm(F.linear(g, n64.fc1.weight, n64.fc1.bias), None, None, n64.bn4.weight, n64.bn4.bias, True))
    g = F.relu(F.batch_norm(F.linear(g, n64.fc2.weight, n64.fc2.bias), None, None, n64.bn5.weight, n64.bn5.bias, True))
    T64 = (F.linear(g, n64.fc3.weight, n64.fc3.bias) + torch.eye(k, device="cuda", dtype=torch.double).view(1, -1)).view(B, k, k)
    (T64 * Wt.double()).sum().backward()
    ref64 = {n: p.grad.clone() for n, p in n64.named_parameters()}
    # mine
    tn = train.TNetTrain(net); grads = train._Grads()
    with torch.no_grad():
        Tm = tn.fwd(rows, k, B, N)
        tn.bwd(Wt.clone(), grads, need_drows=nd)
    print("seed", seed, "need_drows", nd, "T err", float((Tm - T).abs().max()))
    for n, p in net.named_parameters():
        gm = grads.by_param.get(p)
        r = ref[n]
        r64 = ref64[n]
        print(f"   {n:14s} mine-vs-ref32 {float((gm - r).abs().max() / r.abs().max().clamp_min(1e-12)):.2e}  mine-vs-ref64 {float((gm.double() - r64).abs().max() / r64.abs().max().clamp_min(1e-12)):.2e}  ref32-vs-ref64 {float((r.double() - r64).abs().max() / r64.abs().max().clamp_min(1e-12)):.2e}")
