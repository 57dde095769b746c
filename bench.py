#!/usr/bin/env python
"""bench.py — LPD-Net eval embedding throughput (BASELINE.json config C2) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores (oracle port)

One "step" = one pass of the hot path (LPDNet featnet + NetVLAD, eval mode) over one batch of 64 synthetic
4096-point submaps per GPU.  Batch sharding only — no data-path collective ("scaling": "weak").
`value`  = submaps/s with inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    = the same metric through the public bulk-embedding API (evaluate.get_latent_vectors) from pinned HOST
           memory, H2D of every batch and D2H of every descriptor block inside the timed region.
`roofline` = the dominant kernel of the step, timed live with CUDA events around its launches.
`cpu_baseline` = the CPU oracle (a numpy / C restatement of the reference's as-written algorithm) on the box's cores,
           on a bounded sample of the same workload.  The reference itself is Python + torch and lives only in the
           authoring container (/root/reference), so kind == "port".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "LPD-Net eval embedding throughput (featnet=lpdnet, 4096 pts, k=20, NetVLAD K=64 D=1024 -> 256)"
UNIT = "submaps/s"
BATCH, NPTS, KNN = 64, 4096, 20


def ncu_traffic(label):
    """DRAM bytes per launch of a kernel family from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py from dram__bytes_read.sum + dram__bytes_write.sum); None if not captured."""
    try:
        table = json.loads((Path(__file__).resolve().parent / "profiles" / "ncu_traffic.json").read_text())
        return table.get(label, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# algorithmic work per launch of each kernel family, per cloud (min-FLOP decomposition)
def kernel_work(label: str, B: int):
    """-> (flops, bytes) algorithmic per launch, or None"""
    N, k = NPTS, KNN
    if label.startswith("lpd_gemm["):
        M, Nn, K, batch = (int(v) for v in label[9:-1].split("x"))
        return 2.0 * M * Nn * K * batch, 4.0 * batch * (M * K + Nn * K + M * Nn)
    if label.startswith("lpd_gemm_tf32["):
        dims = [int(v) for v in label[14:-1].split("x")]
        M, Nn, K = dims[:3]
        batch = dims[3] if len(dims) > 3 else 1
        return 2.0 * M * Nn * K * batch, 4.0 * batch * (M * K + Nn * K + M * Nn)
    if label.startswith("lpd_knn[C=64") or label.startswith("lpd_knn_tc[C=64"):
        return 2.0 * N * N * 64 * B, 4.0 * B * N * (64 + k)
    if label.startswith("lpd_knn[C=3"):
        return 2.0 * N * N * 3 * B, 4.0 * B * N * (3 + k)
    if label.startswith("lpd_edgeconv_dg[128") or label.startswith("lpd_edgeconv_dg_tf32[128"):
        return 2.0 * N * k * 128 * 128 * B, 4.0 * B * N * (256 + 256 + k)
    if label.startswith("lpd_edge_gather_ext"):
        C = int(label.split("C=")[1].rstrip("]"))
        return 0.0, 4.0 * B * N * (2 * C + C + k)
    if label == "lpd_netvlad_assign":
        return 2.0 * N * 1024 * 64 * B, 4.0 * B * N * (1024 + 64)
    if label == "lpd_pointwise_mlp2":
        return 2.0 * B * N * (3 * 64 + 64 * 64), 4.0 * B * N * (3 + 64)
    if label.startswith("lpd_knn_xyz"):
        return 2.0 * N * N * 3 * B, 4.0 * B * N * (3 + k)
    if label.startswith("lpd_edge_sel_stats") or label.startswith("lpd_edge_bwd_apply") or label.startswith("lpd_edge_bwd_reduce"):
        # train-mode decomposed edge layer: per point the projected rows P, Q (C floats each) and k indices in, the selected
        # row (C floats) + arg (C bytes) out / the gradient rows in and out; the N*k*C edge tensor counts zero
        C = int(label.split("C=")[1].rstrip("]"))
        return 0.0, B * N * (4.0 * (3 * C + k) + C)
    if label.startswith("lpd_edge_materialize") or label.startswith("lpd_edge_sel_dense") or label.startswith("lpd_edge_dense_bwd_apply"):
        C = int(label.split("C=")[1].rstrip("]"))
        return 0.0, 4.0 * B * N * k * C       # the materialised [B*N*k][C] edge rows the DG2 backward needs, once
    return None


def make_roofline(top, tot, cnt, step_ms, peaks, B):
    """roofline object of the dominant kernel family: algorithmic FLOPs (or bytes) of ONE launch / its CUDA-event time."""
    work = kernel_work(top, B)
    if work is None:
        return {"kernel": top, "bound": None, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": ncu_traffic(top),
                "share_of_step": tot[top] / step_ms, "ms_per_launch": tot[top] / cnt[top], "note": "no algorithmic work model for this kernel"}
    flops, byts = work
    per_launch_ms = tot[top] / cnt[top]
    # contractions with >= 64 FLOP per compulsory byte are graded on the tensor pipe; thin GEMMs (K or N <= 128: 30 FLOP/B,
    # far below the ~210 FLOP/B ridge of the measured peaks) and streaming kernels on HBM bandwidth
    if flops > 0 and flops / byts >= 64.0:
        tf = flops / (per_launch_ms * 1e-3) / 1e12
        return {"kernel": top, "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": tf / peaks["bf16_tflops_sustained"], "traffic": ncu_traffic(top),
                "peak_source": peaks["source"] + ", bf16 sustained", "share_of_step": tot[top] / step_ms, "ms_per_launch": per_launch_ms,
                "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": byts,
                "note": "algorithmic FLOPs per launch / CUDA-event time, graded against the dense bf16 tensor peak (an fp16/TF32 "
                        "tensor-core kernel: its own peak is 1x / 0.5x of that; fp32 CUDA-core kernel: ~1/20)"}
    gbs = byts / (per_launch_ms * 1e-3) / 1e9
    return {"kernel": top, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": ncu_traffic(top), "peak_source": peaks["source"] + ", copy bandwidth", "share_of_step": tot[top] / step_ms,
            "ms_per_launch": per_launch_ms, "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": byts,
            "note": "algorithmic (compulsory) bytes per launch / CUDA-event time"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc, self.index = None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_model(device):
    import torch
    from lpdnet_b200 import synth
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    torch.manual_seed(1234)
    model = PointNetVlad(num_points=NPTS, featnet="lpdnet", emb_dims=1024)   # random-init architecture, synthetic weights
    sd = synth.synthetic_state_dict(model)
    model.load_state_dict(sd)
    return model.to(device).eval(), sd


def cpu_baseline(n_clouds: int, threads: int, reps: int = 1):
    """Times the CPU oracle (the reference's as-written forward: SGEMM + top-k kNN, materialised edge tensors,
    un-folded BatchNorm) on `n_clouds` clouds of the workload, one cloud per host thread (numpy's element-wise ops
    are single-threaded, so cloud-level parallelism is what uses all cores).  Returns (submaps/s, seconds)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits
    from lpdnet_b200 import synth
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    from oracle import model_numpy
    shapes = {k: v.shape for k, v in PointNetVlad(num_points=NPTS, featnet="lpdnet", emb_dims=1024).state_dict().items()}
    sd = {k: v.numpy() for k, v in synth.fill_state_dict(shapes).items()}
    x = synth.clouds(n_clouds, NPTS).numpy()
    model_numpy.KNN_IMPL = "blas"      # knn() as written (matmul + top-k); the canonical tie order is for parity only
    best = float("inf")
    try:
        with threadpool_limits(limits=max(1, threads // max(1, min(threads, n_clouds)))):
            for _ in range(reps):
                t0 = time.perf_counter()
                with ThreadPoolExecutor(max_workers=min(threads, n_clouds)) as ex:
                    outs = list(ex.map(lambda b: model_numpy.pointnetvlad_forward(sd, x[b:b + 1], featnet="lpdnet"), range(n_clouds)))
                best = min(best, time.perf_counter() - t0)
    finally:
        model_numpy.KNN_IMPL = "canonical"
    assert np.isfinite(np.concatenate(outs)).all()
    return n_clouds / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = threads
    _, t_full = cpu_baseline(sample, threads)            # warm-up (always one: page-in, BLAS thread pools)
    if args.steps * t_full > 150.0:                       # keep the whole run within a few minutes
        sample = max(1, int(threads * 150.0 / (args.steps * t_full)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_baseline(sample, threads)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: LPD-Net eval embedding, 4096-pt submaps", "step": f"{sample} submaps per step on the host CPU, one per core (bounded sample of the 64-submap batch)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} submaps/step x {args.steps} steps, numpy+C oracle of the reference's as-written forward"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lpdnet_b200 import evaluate, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    ops.set_precision(args.precision)
    model, _ = build_model(device)
    # per-rank shard of the synthetic submap stream: 4 distinct batches so consecutive steps never see the same input
    n_rot = 4
    host = [synth.clouds(BATCH, NPTS, seed=1234 + 97 * rank + i).pin_memory() for i in range(n_rot)]
    dev_in = [h.to(device) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def step(i):
        with torch.no_grad():
            return model(dev_in[i % n_rot])

    for i in range(max(args.warmup, 3)):
        out = step(i)
    barrier()

    # ---- device-resident throughput: K steps, each timed by its own event pair, L2 flushed (untimed) in between ----
    ops.reset_launch_count()
    evs = []
    with ClockSampler(local) as clk:
        barrier()
        for i in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = step(i)
            e.record()
            evs.append((s, e))
        barrier()
    launches = ops.launch_count()
    t_ms = sum(s.elapsed_time(e) for s, e in evs)
    t = torch.tensor([t_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t)
    value = world * BATCH * args.steps / (t_ms * 1e-3)

    # ---- end to end through the public API: pinned host clouds -> descriptors in host memory ----
    big = torch.cat([h[:, 0] for h in host], 0)                      # [4*64, N, 3] host
    reps = max(1, (args.steps + n_rot - 1) // n_rot)
    evaluate.get_latent_vectors(model, big, batch_num=BATCH)   # warm (captures the driver's CUDA graph of a full batch)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n_e2e = 0
    for _ in range(reps):
        desc = evaluate.get_latent_vectors(model, big, batch_num=BATCH)
        n_e2e += big.shape[0]
    e.record()
    barrier()
    te = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e / (float(te) * 1e-3)

    # ---- per-kernel device time of one step (separate, untimed-for-throughput pass) ----
    roofline, breakdown = None, None
    if rank == 0:
        peaks = load_peaks()
        per = {}
        nprof = 3
        for i in range(nprof):
            flush.zero_()
            ops.profile(True)
            step(i)
            rec = ops.profile(False)
            torch.cuda.synchronize(device)
            for label, a, b in rec:
                per.setdefault(label, []).append(a.elapsed_time(b))
        tot = {k_: sum(v) / nprof for k_, v in per.items()}                 # ms per step per kernel family
        cnt = {k_: len(v) / nprof for k_, v in per.items()}
        step_ms = sum(tot.values())
        breakdown = {k_: {"ms_per_step": round(v, 4), "share": round(v / step_ms, 4), "launches": cnt[k_]}
                     for k_, v in sorted(tot.items(), key=lambda kv: -kv[1])}
        top = max(tot, key=tot.get)
        roofline = make_roofline(top, tot, cnt, step_ms, peaks, BATCH)

    if rank == 0:
        threads = os.cpu_count() or 1
        cpu_v, cpu_t = cpu_baseline(threads, threads)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"precision": ("tf32 tensor-core GEMMs (fp32 storage, fp32 accumulate; kNN and its input layers exact fp32)"
                                         if args.precision == "tf32" else "strict fp32 FFMA"),
                           "workload": "C2: LPD-Net eval embedding (featnet=lpdnet, kNN k=20 graph features + NetVLAD K=64 D=1024 -> 256)",
                           "submaps_per_gpu_per_step": BATCH, "points": NPTS, "sharding": f"batch-sharded dp{world}, no collective",
                           "l2": "256 MiB memset between timed steps (untimed); 4 rotating input batches; intermediates > 1 GiB/step"},
                "clocks": clk.result, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * NPTS * 3 * 4, "d2h_bytes_per_step": BATCH * 256 * 4,
                        "api": "lpdnet_b200.evaluate.get_latent_vectors (pinned host clouds -> host descriptors)"},
                "roofline": roofline, "kernel_breakdown": breakdown,
                "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{threads} submaps (of the 64-submap batch) in {cpu_t:.1f} s, one per core, numpy oracle of the reference's as-written forward"}}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ======================================================================================================================
# C3: LPD-Net training step (train-mode forward, lazy quadruplet loss, backward, gradient all-reduce, Adam)
# ======================================================================================================================
TRAIN_METRIC = "LPD-Net training-step throughput (featnet=lpdnet, 4096 pts, lazy quadruplet loss, batch_num_queries=2 per GPU, Adam)"
TRAIN_BQ, TRAIN_P, TRAIN_NN = 2, 2, 18
TRAIN_CLOUDS = TRAIN_BQ * (1 + TRAIN_P + TRAIN_NN + 1)      # 44 clouds per GPU per step


def synth_tuples(seed: int):
    """one batch of training tuples in the reference's DataLoader layout (host tensors)"""
    import torch
    from lpdnet_b200 import synth
    x = synth.clouds(TRAIN_CLOUDS, NPTS, seed=seed).view(TRAIN_BQ, 1 + TRAIN_P + TRAIN_NN + 1, NPTS, 3)
    return tuple(t.contiguous() for t in torch.split(x, [1, TRAIN_P, TRAIN_NN, 1], dim=1))


def cpu_baseline_train(n_tuples: int, threads: int):
    """Times the differentiable CPU oracle (oracle/model_torch.py: the reference's as-written forward + torch autograd +
    the same loss) on `n_tuples` tuples of 22 clouds.  Returns (submaps/s, seconds)."""
    import torch
    from lpdnet_b200 import synth
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    from oracle import model_torch
    torch.set_num_threads(threads)
    shapes = {k: v.shape for k, v in PointNetVlad(num_points=NPTS, featnet="lpdnet", emb_dims=1024).state_dict().items()}
    sd = synth.fill_state_dict(shapes)
    x = synth.clouds(22 * n_tuples, NPTS)
    t0 = time.perf_counter()
    _, loss, grads, _ = model_torch.train_step(sd, x, n_tuples)
    dt = time.perf_counter() - t0
    assert bool(torch.isfinite(loss)) and len(grads) > 20
    return 22 * n_tuples / dt, dt


def run_reference_train(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        cpu_baseline_train(1, threads)
        n += 22
    dt = time.perf_counter() - t0
    v = n / dt
    line = {"impl": "reference", "metric": TRAIN_METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: LPD-Net training step, 4096-pt submaps", "step": "1 tuple (22 submaps) per step on the host CPU: forward + loss + autograd backward (bounded sample of the 44-submap step; no optimizer step)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"22 submaps/step x {args.steps} steps, torch-CPU oracle of the reference's as-written training forward/backward"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_train(args):
    import torch
    import torch.distributed as dist
    from lpdnet_b200 import ops, optim, synth
    from lpdnet_b200 import train_pointnetvlad as TP
    from lpdnet_b200.util.PointNetVlad import PointNetVlad

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    ops.set_precision(args.precision)
    torch.manual_seed(1234)
    model = PointNetVlad(num_points=NPTS, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))          # identical replicas on every rank
    model = model.to(device).train()
    # lr 1e-7: with the reference's default 1e-3 (or even 1e-5) Adam drives the hinge loss of these 4 synthetic tuple batches
    # to exactly 0 within ~10 steps and every gradient becomes zero; the kernels do the same work either way (nothing is
    # data-dependent), but a tiny lr keeps the loss and gradients non-trivial for the whole timed region.
    opt = optim.Adam(model.parameters(), lr=1e-7)
    n_rot = 4
    host = [tuple(t.pin_memory() for t in synth_tuples(4321 + 131 * rank + i)) for i in range(n_rot)]
    dev_in = [tuple(t.to(device) for t in h) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def step(batch):
        return TP.train_step(model, opt, *batch, margin_1=0.5, margin_2=0.2)

    for i in range(max(args.warmup, 3)):
        loss = step(dev_in[i % n_rot])
    barrier()
    ops.reset_launch_count()
    evs = []
    with ClockSampler(local) as clk:
        barrier()
        for i in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            loss = step(dev_in[i % n_rot])
            e.record()
            evs.append((s, e))
        barrier()
    launches = ops.launch_count()
    t = torch.tensor([sum(s.elapsed_time(e) for s, e in evs)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t)
    value = world * TRAIN_CLOUDS * args.steps / (t_ms * 1e-3)
    last_loss = float(loss)

    # ---- end to end: pinned host tuples -> H2D -> step -> loss value back on the host, every step ----
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        loss_host = float(step(host[i % n_rot]))
    e.record()
    barrier()
    te = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * TRAIN_CLOUDS * args.steps / (float(te) * 1e-3)

    roofline, breakdown = None, None
    # per-kernel device time of a step: every rank runs these steps (the gradient all-reduce inside optimizer.step() is a
    # collective), only rank 0 records
    per = {}
    nprof = 2
    for i in range(nprof):
        flush.zero_()
        ops.profile(rank == 0)
        step(dev_in[i % n_rot])
        rec = ops.profile(False)
        torch.cuda.synchronize(device)
        for label, a, b in (rec or []):
            per.setdefault(label, []).append(a.elapsed_time(b))
    if rank == 0:
        peaks = load_peaks()
        tot = {k_: sum(v) / nprof for k_, v in per.items()}
        cnt = {k_: len(v) / nprof for k_, v in per.items()}
        step_ms = sum(tot.values())
        breakdown = {k_: {"ms_per_step": round(v, 4), "share": round(v / step_ms, 4), "launches": cnt[k_]}
                     for k_, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]}
        top = max(tot, key=tot.get)
        roofline = make_roofline(top, tot, cnt, step_ms, peaks, TRAIN_CLOUDS)
        threads = os.cpu_count() or 1
        cpu_v, cpu_t = cpu_baseline_train(1, threads)
        nparams = sum(p.numel() for p in model.parameters())
        line = {"metric": TRAIN_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": "C3: LPD-Net training step (train-mode forward, lazy quadruplet loss m1=0.5 m2=0.2, backward, "
                                       "gradient all-reduce, fused Adam)",
                           "precision": args.precision, "submaps_per_gpu_per_step": TRAIN_CLOUDS, "tuples_per_gpu": TRAIN_BQ, "points": NPTS,
                           "sharding": f"whole tuples per GPU (dp{world}), per-rank BatchNorm statistics, one NCCL all-reduce of the flat fp32 gradient buffer ({4 * nparams / 1e6:.1f} MB) per step",
                           "l2": "256 MiB memset between timed steps (untimed); 4 rotating tuple batches", "last_loss": last_loss},
                "clocks": clk.result, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": TRAIN_CLOUDS * NPTS * 3 * 4, "d2h_bytes_per_step": 4,
                        "api": "lpdnet_b200.train_pointnetvlad.train_step (pinned host tuples -> loss value on the host)", "last_loss": loss_host},
                "roofline": roofline, "kernel_breakdown": breakdown,
                "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"1 tuple (22 submaps) forward+loss+backward in {cpu_t:.1f} s, torch-CPU oracle of the reference's as-written training path"}}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) print to stdout; the contract is ONE JSON line there.  Point fd 1 at
    stderr for the whole run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"],
                    help="tf32: dense layers on tcgen05 tensor cores (descriptor error vs the reference measured <= 1e-4); "
                         "fp32: every layer in strict fp32 FFMA arithmetic")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 (default, the configuration BASELINE.json's metric is quoted on): eval embedding of 64 submaps per GPU; "
                         "c3: training step on 2 tuples = 44 submaps per GPU")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        (run_reference_train if args.workload == "c3" else run_reference)(args)
    elif args.workload == "c3":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
