#!/usr/bin/env python
"""bench.py — the LPD-Net hot path on N B200s of one node: BASELINE.json's metric (C2 eval embedding, submaps/s) as the
top-level line, and the other configurations of BASELINE.json under `workloads` (C3 training step, C4 retrieval, C5 stress).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation on the host cores
    python bench.py --workload c3|c4|c5 ...                  # one workload as the top-level line (development)

A "step" is one pass of the hot path over one batch of synthetic input per GPU:
  C2  64 submaps x 4096 points through PointNetVlad(featnet=lpdnet).eval()             batch-sharded, no collective ("weak")
  C3  2 tuples = 44 submaps: train-mode forward, lazy quadruplet loss, backward, NCCL gradient all-reduce, fused Adam ("weak")
  C4  recall@N of 23 runs x 956 descriptors (506 run pairs, 66,792 searches): database runs sharded over the ranks, counters
      all-reduced; plus the one-big-database search (21,988 rows sharded, NCCL all-gather + top-k merge)          ("strong")
  C5  32 clouds x 16384 points, k = 32 (one GPU's share of the 256-cloud stress batch)                              ("weak")
`value`  = units / s with inputs resident in HBM (CUDA events around every step, max over ranks, L2 flushed between steps).
`e2e`    = the same through the public API from pinned HOST buffers: H2D of every input and D2H of every result in the region.
`roofline` = the dominant kernel family, timed live with CUDA events; `rooflines` = every family of the step.
`cpu_baseline` = the UNMODIFIED reference (oracle/_ref: staged copy of its modules, kind "reference"; the numpy/C port if that
           copy is absent, kind "port") on the box's host cores, bounded sample, rank 0 at N = 1 only.
At N > 1 the line carries `selfcheck`: the all-reduced gradient equals the sum of the per-rank gradients, and the merged
top-25 of the row-sharded database equals the unsharded search.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "LPD-Net eval embedding throughput (featnet=lpdnet, 4096 pts, k=20, NetVLAD K=64 D=1024 -> 256)"
UNIT = "submaps/s"
BATCH, NPTS, KNN = 64, 4096, 20
C5_BATCH, C5_NPTS, C5_KNN = 32, 16384, 32
TRAIN_METRIC = "LPD-Net training-step throughput (featnet=lpdnet, 4096 pts, lazy quadruplet loss, batch_num_queries=2 per GPU, Adam)"
TRAIN_BQ, TRAIN_P, TRAIN_NN = 2, 2, 18
TRAIN_CLOUDS = TRAIN_BQ * (1 + TRAIN_P + TRAIN_NN + 1)      # 44 clouds per GPU per step
C4_METRIC = "recall@N retrieval throughput (23 runs x 956 x 256-d descriptors, 506 run pairs, 25-NN per query and database run)"
C5_METRIC = "LPD-Net eval embedding throughput, stress shape (16384 pts, k=32)"


# ======================================================================================================================
# shared plumbing
# ======================================================================================================================
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def measure_tf32_peak(device):
    """dense TF32 matmul peak of this GPU, measured the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, best of 10,
    CUDA events) — the denominator that actually applies to the kind::tf32 kernels.  Library call: measurement only."""
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=device)
        b = torch.randn(n, n, device=device)
        for _ in range(3):
            a @ b
        best = float("inf")
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            a @ b
            e.record()
            torch.cuda.synchronize(device)
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def ncu_traffic(label):
    """DRAM bytes per launch of a kernel family from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py from dram__bytes_read.sum + dram__bytes_write.sum); None if not captured."""
    try:
        table = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        return table.get(label, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def kernel_work(label: str, B: int, N: int = NPTS, k: int = KNN):
    """-> (flops, bytes) ALGORITHMIC per launch (min-FLOP decomposition, compulsory bytes), or None"""
    if label.startswith("lpd_gemm["):
        M, Nn, K, batch = (int(v) for v in label[9:-1].split("x"))
        return 2.0 * M * Nn * K * batch, 4.0 * batch * (M * K + Nn * K + M * Nn)
    if label.startswith(("lpd_gemm_f16[", "lpd_gemm_f16_tn[")):
        dims = [int(v) for v in label[label.index("[") + 1:-1].split("x")]
        M, Nn, K = dims[:3]
        batch = dims[3] if len(dims) > 3 else 1
        return 2.0 * M * Nn * K * batch, 2.0 * batch * (M * K + Nn * K) + 4.0 * batch * M * Nn     # fp16 operands
    if label.startswith("lpd_gemm_tf32[") or label.startswith("lpd_gemm_tf32_tn["):
        dims = [int(v) for v in label[label.index("[") + 1:-1].split("x")]
        M, Nn, K = dims[:3]
        batch = dims[3] if len(dims) > 3 else 1
        return 2.0 * M * Nn * K * batch, 4.0 * batch * (M * K + Nn * K + M * Nn)
    if label.startswith("lpd_knn[C=64") or label.startswith("lpd_knn_tc[C=64"):
        return 2.0 * N * N * 64 * B, 4.0 * B * N * (64 + k)
    if label.startswith("lpd_knn[C=3") or label.startswith("lpd_knn_xyz"):
        return 2.0 * N * N * 3 * B, 4.0 * B * N * (3 + k)
    if label.startswith("lpd_edgeconv_dg_f16[128"):
        return 2.0 * N * k * 128 * 128 * B, 2.0 * B * N * (256 + 256) + 4.0 * B * N * k
    if label.startswith("lpd_edgeconv_dg[128") or label.startswith("lpd_edgeconv_dg_tf32[128"):
        return 2.0 * N * k * 128 * 128 * B, 4.0 * B * N * (256 + 256 + k)
    if label.startswith("lpd_edge_gather_ext"):
        C = int(label.split("C=")[1].rstrip("]"))
        return 0.0, 4.0 * B * N * (2 * C + C + k)
    if label.startswith("lpd_gemm_softmax64["):
        M, Nn, K = (int(v) for v in label[label.index("[") + 1:-1].split("x"))
        return 2.0 * M * Nn * K, 2.0 * (M * K + Nn * K + M * Nn)                                  # fp16 in, fp16 assignment out
    if label == "lpd_netvlad_assign":
        return 2.0 * N * 1024 * 64 * B, 4.0 * B * N * (1024 + 64)
    if label.startswith("lpd_conv3_vlad"):
        # fused conv3 512->1024 + BN/act + NetVLAD assign + softmax + aggregate: the 1024-d feature map never reaches HBM
        return 2.0 * B * N * (512 * 1024 + 2 * 1024 * 64), 4.0 * B * (N * 512 + 1024 * 64)
    if label == "lpd_pointwise_mlp2":
        return 2.0 * B * N * (3 * 64 + 64 * 64), 4.0 * B * N * (3 + 64)
    if label.startswith("lpd_edge_sel_stats") or label.startswith("lpd_edge_bwd_apply") or label.startswith("lpd_edge_bwd_reduce"):
        C = int(label.split("C=")[1].rstrip("]"))
        return 0.0, B * N * (4.0 * (3 * C + k) + C)
    if label.startswith("lpd_edge_materialize") or label.startswith("lpd_edge_sel_dense") or label.startswith("lpd_edge_dense_bwd_apply"):
        C = int(label.split("C=")[1].rstrip("]"))
        return 0.0, 4.0 * B * N * k * C
    if label.startswith("lpd_retrieval_tc["):
        Nq, Ndb, D = (int(v) for v in label[label.index("[") + 1:-1].split("x"))
        return 2.0 * Nq * Ndb * D, 4.0 * (Nq * D + Ndb * D)
    return None


def roofline_of(label, ms_total, launches, step_ms, peaks, B, N=NPTS, k=KNN):
    """roofline object of one kernel family: algorithmic FLOPs (or bytes) of ONE launch / its CUDA-event time"""
    work = kernel_work(label, B, N, k)
    per_launch_ms = ms_total / launches
    base = {"kernel": label, "share_of_step": ms_total / step_ms, "ms_per_launch": per_launch_ms, "traffic": ncu_traffic(label)}
    if work is None:
        return {**base, "bound": None, "achieved": None, "peak": None, "unit": None, "frac": None,
                "note": "no algorithmic work model for this kernel"}
    flops, byts = work
    base.update({"algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": byts})
    # contractions with >= 64 FLOP per compulsory byte are graded on the tensor pipe; thin GEMMs (K or N <= 128: 30 FLOP/B,
    # far below the ~210 FLOP/B ridge of the measured peaks) and streaming kernels on HBM bandwidth
    if flops > 0 and flops / byts >= 64.0:
        tf = flops / (per_launch_ms * 1e-3) / 1e12
        out = {**base, "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
               "frac": tf / peaks["bf16_tflops_sustained"], "peak_source": peaks["source"] + ", bf16 sustained"}
        if peaks.get("tf32_tflops"):
            out["frac_of_tf32_peak"] = tf / peaks["tf32_tflops"]
        return out
    gbs = byts / (per_launch_ms * 1e-3) / 1e9
    return {**base, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "peak_source": peaks["source"] + ", copy bandwidth"}


class ClockSampler:
    """SM clock, power, temperature and throttle reasons of this rank's GPU during the timed region, sampled in-process through
    NVML from a thread (round 1 forked one `nvidia-smi -lms 20` per rank: eight concurrent 50 Hz driver queries)."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index: int, enable: bool = True, period_s: float = 0.01):
        self.index, self.enable, self.period = index, enable, period_s
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            sm, reasons, power, temp = [], set(), [], []
            while True:
                sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    temp.append(pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU))
                except Exception:  # noqa: BLE001
                    pass
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        reasons.add(name)
                if self._stop.wait(self.period):
                    break
            self.result = {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": float(mx), "reasons": sorted(reasons),
                           "samples": len(sm), "power_w_max": max(power) if power else None, "temp_c_max": max(temp) if temp else None,
                           "how": "NVML in-process thread"}
        except Exception as ex:  # noqa: BLE001 — NVML missing / not permitted: fall back to one nvidia-smi query
            self.result = self._smi_once(str(ex))

    def _smi_once(self, why):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().split(",")]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")],
                    "samples": 1, "how": f"nvidia-smi once after the region (NVML unavailable: {why[:80]})"}
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}

    def __enter__(self):
        if self.enable:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=5)


class Ctx:
    """one process per GPU: rank / device / barrier / max-over-ranks"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU reference")
        self.pin_cores()
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)   # > 126 MB L2
        self.peaks = load_peaks()
        self.steps, self.warmup = args.steps, max(args.warmup, 3)

    def pin_cores(self):
        """give every local rank its own slice of the allowed cores (round 1: all 8 ranks shared cores 0-31 of NUMA node 0 and
        the max-over-ranks step time grew 12 % at 8 GPUs with identical per-kernel times)"""
        try:
            cores = sorted(os.sched_getaffinity(0))
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)))
            per = len(cores) // max(1, lw)
            if lw > 1 and per >= 1:
                os.sched_setaffinity(0, set(cores[self.local * per:(self.local + 1) * per]))
            self.cores = sorted(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            self.cores = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def flush(self):
        self.flush_buf.zero_()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], device=self.device, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def gather_floats(self, v: float):
        t = self.torch.tensor([v], device=self.device, dtype=self.torch.float64)
        if self.world == 1:
            return [float(t)]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(x) for x in out]

    def timed_steps(self, step, steps=None):
        """K steps, each bracketed by its own CUDA-event pair on the launching stream, L2 flushed (untimed) in between;
        barrier + synchronize on both sides.  -> (ms summed over the steps, max over ranks; per-rank sums; per-step list)"""
        torch = self.torch
        steps = steps or self.steps
        evs = []
        self.barrier()
        for i in range(steps):
            self.flush()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step(i)
            e.record()
            evs.append((s, e))
        self.barrier()
        per_step = [s.elapsed_time(e) for s, e in evs]
        mine = sum(per_step)
        return self.max_over_ranks(mine), self.gather_floats(mine), per_step

    def rank_report(self, per_step, clk):
        """per-rank diagnostics gathered on rank 0: [min, median, max] step time and the GPU's clocks / power during the region.
        A rank whose MINIMUM step time is high runs on a slower GPU (clocks, power cap); a high maximum alone is a straggling step."""
        mine = {"rank": self.rank, "gpu": self.local, "step_ms_min_median_max": [round(min(per_step), 4), round(statistics.median(per_step), 4),
                                                                                 round(max(per_step), 4)],
                "step_ms": [round(v, 3) for v in per_step],
                "sm_mhz_median": clk.get("sm_mhz"), "sm_mhz_min": clk.get("sm_min_mhz"), "power_w_max": clk.get("power_w_max"),
                "temp_c_max": clk.get("temp_c_max"), "reasons": clk.get("reasons")}
        if self.world == 1:
            return [mine]
        out = [None] * self.world
        self.dist.all_gather_object(out, mine)
        return out

    def timed_region(self, fn):
        """one CUDA-event pair around fn() (end-to-end regions that synchronise inside) -> ms, max over ranks"""
        torch = self.torch
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        self.barrier()
        return self.max_over_ranks(s.elapsed_time(e))

    def profile_step(self, step, nprof, record=True):
        """per-kernel-family device time of a step (separate eager pass, one CUDA-event pair per C-ABI call)"""
        from lpdnet_b200 import ops
        per = {}
        for i in range(nprof):
            self.flush()
            ops.profile(record)
            step(i)
            rec = ops.profile(False)
            self.torch.cuda.synchronize(self.device)
            for label, a, b in (rec or []):
                per.setdefault(label, []).append(a.elapsed_time(b))
        tot = {k_: sum(v) / nprof for k_, v in per.items()}
        cnt = {k_: len(v) / nprof for k_, v in per.items()}
        return tot, cnt


class GraphStep:
    """model(x) for a fixed input shape captured into ONE CUDA graph: a timed step is a device-to-device copy of the step's
    input into the graph's static buffer + one replay, so the 8 ranks of a node do not compete for host cores to issue ~25
    launches per step each (the kernels and their order are exactly those of the eager call)."""

    def __init__(self, ctx, model, example):
        from lpdnet_b200 import ops
        torch = ctx.torch
        self.x = torch.zeros_like(example)
        side = torch.cuda.Stream(device=ctx.device)
        side.wait_stream(torch.cuda.current_stream(ctx.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                model(self.x)
            ops.reset_launch_count()
            model(self.x)
            self.launches = ops.launch_count()
        torch.cuda.current_stream(ctx.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.y = model(self.x)

    def __call__(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y


def reference_modules():
    """(L, PNV, RL) of the UNMODIFIED reference staged under oracle/_ref (or /root/reference when present), else None"""
    try:
        from oracle import ref_loader
        if ref_loader.reference_root() is None:
            return None
        return ref_loader.import_reference()
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] reference modules unavailable ({ex}); falling back to the numpy/C port", file=sys.stderr)
        return None


def build_model(device, npts=NPTS, k=None):
    import torch
    from lpdnet_b200 import synth
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    torch.manual_seed(1234)
    model = PointNetVlad(num_points=npts, featnet="lpdnet", emb_dims=1024)   # random-init architecture, synthetic weights
    model.load_state_dict(synth.synthetic_state_dict(model))
    if k is not None:
        model.emb_nn.k = k           # reference lpdnet_model.py:156: k is a mutable attribute, not a constructor argument
    return model.to(device).eval()


# ======================================================================================================================
# CPU arms: the reference's own modules on the host cores
# ======================================================================================================================
def _ref_model(PNV, npts, k=None, train=False):
    import torch
    from lpdnet_b200 import synth
    torch.manual_seed(1234)
    model = PNV.PointNetVlad(num_points=npts, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))
    if k is not None:
        model.emb_nn.k = k
    return model.train(train)


def cpu_embed(n_clouds: int, threads: int, chunk: int = 16, npts: int = NPTS, k=None, model=None):
    """eval embedding of n_clouds synthetic submaps on the host: the reference's PointNetVlad on `threads` torch threads in
    chunks of `chunk` clouds (a 16 x 4096 chunk peaks at 4.7 GB), or — without oracle/_ref — the numpy port, one cloud per
    thread.  -> (submaps/s, seconds, kind, model)"""
    import numpy as np
    import torch
    from lpdnet_b200 import synth
    ref = reference_modules()
    x = synth.clouds(n_clouds, npts)
    if ref is not None:
        torch.set_num_threads(threads)
        model = model or _ref_model(ref[1], npts, k)
        t0 = time.perf_counter()
        outs = []
        with torch.no_grad():
            for lo in range(0, n_clouds, chunk):
                outs.append(model(x[lo:lo + chunk]))
        dt = time.perf_counter() - t0
        assert bool(torch.isfinite(torch.cat(outs)).all())
        return n_clouds / dt, dt, "reference", model
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    from oracle import model_numpy
    shapes = {k_: v.shape for k_, v in PointNetVlad(num_points=npts, featnet="lpdnet", emb_dims=1024).state_dict().items()}
    sd = {k_: v.numpy() for k_, v in synth.fill_state_dict(shapes).items()}
    xn = x.numpy()
    model_numpy.KNN_IMPL = "blas"
    try:
        with threadpool_limits(limits=max(1, threads // max(1, min(threads, n_clouds)))):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=min(threads, n_clouds)) as ex:
                outs = list(ex.map(lambda b: model_numpy.pointnetvlad_forward(sd, xn[b:b + 1], featnet="lpdnet", **({"k": k} if k else {})),
                                   range(n_clouds)))
            dt = time.perf_counter() - t0
    finally:
        model_numpy.KNN_IMPL = "canonical"
    assert np.isfinite(np.concatenate(outs)).all()
    return n_clouds / dt, dt, "port", None


def cpu_train_step(n_tuples: int, threads: int):
    """one training step of the reference on the host (train-mode forward, lazy quadruplet loss, autograd backward, Adam) on
    n_tuples tuples of 22 clouds (15 GB per tuple).  -> (submaps/s, seconds, kind)"""
    import torch
    from lpdnet_b200 import synth
    ref = reference_modules()
    torch.set_num_threads(threads)
    x = synth.clouds(22 * n_tuples, NPTS)
    if ref is not None:
        _, PNV, RL = ref
        model = _ref_model(PNV, NPTS, train=True)
        opt = torch.optim.Adam(model.parameters(), lr=1e-7)
        t0 = time.perf_counter()
        opt.zero_grad()
        out = model(x).view(n_tuples, -1, 256)
        q, pos, neg, other = torch.split(out, [1, TRAIN_P, TRAIN_NN, 1], dim=1)
        loss = RL.quadruplet_loss(q, pos, neg, other, 0.5, 0.2, use_min=True, lazy=True, ignore_zero_loss=False)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        assert bool(torch.isfinite(loss))
        return 22 * n_tuples / dt, dt, "reference"
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    from oracle import model_torch
    shapes = {k_: v.shape for k_, v in PointNetVlad(num_points=NPTS, featnet="lpdnet", emb_dims=1024).state_dict().items()}
    sd = synth.fill_state_dict(shapes)
    t0 = time.perf_counter()
    _, loss, grads, _ = model_torch.train_step(sd, x, n_tuples)
    dt = time.perf_counter() - t0
    assert bool(torch.isfinite(loss)) and len(grads) > 20
    return 22 * n_tuples / dt, dt, "port"


def cpu_recall(n_pairs: int):
    """the reference's get_recall (sklearn KDTree build + one query(k=25) per query, evaluate.py:162-206; single-threaded as
    written) on the first n_pairs ordered run pairs of the synthetic evaluation set.  -> (searches/s, seconds, kind, searches)"""
    from lpdnet_b200 import synth
    DB, Q, SETS = synth.descriptor_database()
    try:
        from oracle import ref_loader
        get_recall, kind = ref_loader.extract_get_recall(), "reference"
    except Exception:  # noqa: BLE001
        from oracle import recall_numpy
        get_recall, kind = recall_numpy.get_recall, "port"
    pairs = [(m, n) for m in range(len(DB)) for n in range(len(DB)) if m != n][:n_pairs]
    t0 = time.perf_counter()
    searches = 0
    for m, n in pairs:
        get_recall(m, n, DB, Q, SETS)
        searches += len(Q[n])
    dt = time.perf_counter() - t0
    return searches / dt, dt, kind, searches


# ======================================================================================================================
# C2 / C5: eval embedding
# ======================================================================================================================
FAMILIES = {   # kernel-label prefix -> family of the step (SURVEY §8d stages)
    "feature kNN": ("lpd_knn_tc", "lpd_knn[C=64"),
    "xyz kNN": ("lpd_knn_xyz", "lpd_knn[C=3"),
    "EdgeConv DG1+DG2": ("lpd_edgeconv_dg",),
    "EdgeConv SN1 gather": ("lpd_edge_gather_ext",),
    "NetVLAD (assign, aggregate, norms, hidden, gating)": ("lpd_netvlad", "lpd_softmax64", "lpd_splitk_reduce", "lpd_gemm_tf32_tn[1024x64",
                                                           "lpd_gemm_f16_tn[1024x64", "lpd_gemm[", "lpd_conv3_vlad", "lpd_hidden", "lpd_gemm_softmax64", "lpd_gemm_f16[262144x64", "lpd_gemm_tf32[262144x64",
                                                           "lpd_gemm_tf32_tn[64x256", "lpd_transpose_split3"),
}


def descriptor_error_vs_reference(ctx, model):
    """max-abs error of this precision mode's descriptors against the UNMODIFIED reference's fp32 CPU output on the committed
    golden case (tests/golden/c2_lpdnet_eval.npz: 4 x 4096-point clouds, same synthetic weights)"""
    try:
        import numpy as np
        from lpdnet_b200 import synth
        g = np.load(ROOT / "tests" / "golden" / "c2_lpdnet_eval.npz", allow_pickle=False)
        with ctx.torch.no_grad():
            out = model(synth.clouds(4, NPTS).to(ctx.device)).cpu().numpy()
        return float(np.abs(out - g["out"]).max())
    except Exception as ex:  # noqa: BLE001
        return f"unavailable: {ex}"


def bench_embed(ctx, tag: str):
    """C2 (tag 'c2') or C5 (tag 'c5') -> result dict (complete on rank 0)"""
    import torch
    from lpdnet_b200 import evaluate, ops, synth
    args = ctx.args
    B, N, k = (BATCH, NPTS, KNN) if tag == "c2" else (C5_BATCH, C5_NPTS, C5_KNN)
    ops.set_precision(args.precision)
    model = build_model(ctx.device, N, None if tag == "c2" else k)
    n_rot = 4
    host = [synth.clouds(B, N, seed=1234 + 97 * ctx.rank + i) for i in range(n_rot)]
    dev_in = [h.to(ctx.device) for h in host]
    step_graph = GraphStep(ctx, model, dev_in[0])

    def step(i):
        return step_graph(dev_in[i % n_rot])

    def step_eager(i):
        with torch.no_grad():
            return model(dev_in[i % n_rot])

    for i in range(ctx.warmup):
        step(i)
    with ClockSampler(ctx.local) as clk:
        t_ms, per_rank, per_step = ctx.timed_steps(step)
    value = ctx.world * B * ctx.steps / (t_ms * 1e-3)
    launches = step_graph.launches * ctx.steps
    ranks = ctx.rank_report(per_step, clk.result)

    # ---- end to end through the public API: PINNED host clouds -> H2D -> descriptors -> D2H into host memory ----
    # one call embeds 16 batches (1,024 submaps at C2: the size of one run of the reference's evaluation sets, evaluate.py:96-159)
    e2e_batches = 4 * n_rot if tag == "c2" else n_rot
    big = torch.cat([h[:, 0] for h in host] * (e2e_batches // n_rot), 0).pin_memory()   # [e2e_batches * B, N, 3]
    reps = max(1, (ctx.steps + e2e_batches - 1) // e2e_batches)
    evaluate.get_latent_vectors(model, big, batch_num=B)                 # warm (stages buffers, captures the driver's own graph)
    te = ctx.timed_region(lambda: [evaluate.get_latent_vectors(model, big, batch_num=B) for _ in range(reps)])
    e2e_value = ctx.world * reps * big.shape[0] / (te * 1e-3)

    res = None
    tot, cnt = ctx.profile_step(step_eager, 3, record=ctx.rank == 0)
    if ctx.rank == 0:
        step_ms = sum(tot.values())
        breakdown = {k_: {"ms_per_step": round(v, 4), "share": round(v / step_ms, 4), "launches": cnt[k_]}
                     for k_, v in sorted(tot.items(), key=lambda kv: -kv[1])}
        fams = {}
        for name, prefixes in FAMILIES.items():
            labels = [l for l in tot if l.startswith(prefixes)]
            if labels:
                fams[name] = {"ms_per_step": round(sum(tot[l] for l in labels), 4), "share": round(sum(tot[l] for l in labels) / step_ms, 4),
                              "kernels": labels}
        top = max(tot, key=tot.get)
        rooflines = [roofline_of(l, tot[l], cnt[l], step_ms, ctx.peaks, B, N, k) for l in sorted(tot, key=tot.get, reverse=True)
                     if tot[l] / step_ms >= 0.02]
        # the metric's named subset: feature-space kNN + xyz kNN + NetVLAD, min-FLOPs of SURVEY §8d over their summed time
        named = [l for l in tot if l.startswith(FAMILIES["feature kNN"] + FAMILIES["xyz kNN"] + FAMILIES["NetVLAD (assign, aggregate, norms, hidden, gating)"])]
        named_ms = sum(tot[l] for l in named)
        named_flops = B * (2.0 * N * N * 64 + 2.0 * N * N * 3 + 2 * 2.0 * N * 1024 * 64) + 2.0 * B * 65536 * 256
        knn_netvlad = {"kernels": named, "ms_per_step": named_ms, "algorithmic_flops": named_flops,
                       "achieved_tflops": named_flops / (named_ms * 1e-3) / 1e12,
                       "frac_of_bf16_sustained": named_flops / (named_ms * 1e-3) / 1e12 / ctx.peaks["bf16_tflops_sustained"]}
        res = {"metric": METRIC if tag == "c2" else C5_METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": ctx.steps,
               "warmup": ctx.warmup, "ms_per_step": t_ms / ctx.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": {"f16": "f16", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
               "config": {"workload": ("C2: LPD-Net eval embedding (featnet=lpdnet, kNN k=20 graph features + NetVLAD K=64 D=1024 -> 256), 64 x 4096-pt submaps per GPU per step"
                                       if tag == "c2" else
                                       "C5: LPD-Net eval embedding, stress shape: 32 clouds x 16384 pts per GPU per step (one GPU's share of the 256-cloud batch at 8 GPUs), k=32"),
                          "precision": {"f16": "fp16 activations and operands downstream of the kNN (round to nearest), tcgen05 kind::f16, fp32 accumulate; "
                                               "kNN and its input layers exact fp32",
                                        "tf32": "tf32 tensor-core GEMMs (fp32 storage, fp32 accumulate; kNN and its input layers exact fp32)",
                                        "fp32": "strict fp32"}[args.precision],
                          "submaps_per_gpu_per_step": B, "points": N, "k": k, "sharding": f"batch-sharded dp{ctx.world}, no collective",
                          "launch": "each timed step = D2D copy of the step's input + ONE CUDA-graph replay of the eager kernel sequence",
                          "l2": "256 MiB memset between timed steps (untimed); 4 rotating input batches; intermediates > 1 GiB/step"},
               "clocks": clk.result, "gpu_launches": launches,
               "per_rank_ms_per_step": [round(v / ctx.steps, 4) for v in per_rank], "ranks": ranks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * N * 3 * 4, "d2h_bytes_per_step": B * 256 * 4,
                       "api": "lpdnet_b200.evaluate.get_latent_vectors (pinned host clouds -> H2D -> graph replay -> D2H host descriptors)",
                       "submaps_per_call": int(big.shape[0]), "calls": reps},
               "roofline": roofline_of(top, tot[top], cnt[top], step_ms, ctx.peaks, B, N, k), "rooflines": rooflines,
               "knn_netvlad": knn_netvlad, "families": fams, "kernel_breakdown": breakdown}
        if tag == "c2":
            res["config"]["descriptor_max_abs_err_vs_reference"] = descriptor_error_vs_reference(ctx, model)
    del step_graph, model              # the model owns the embedding driver's staging ring and captured graph
    torch.cuda.empty_cache()
    return res


# ======================================================================================================================
# C3: training step
# ======================================================================================================================
def synth_tuples(seed: int):
    """one batch of training tuples in the reference's DataLoader layout (host tensors)"""
    import torch
    from lpdnet_b200 import synth
    x = synth.clouds(TRAIN_CLOUDS, NPTS, seed=seed).view(TRAIN_BQ, 1 + TRAIN_P + TRAIN_NN + 1, NPTS, 3)
    return tuple(t.contiguous() for t in torch.split(x, [1, TRAIN_P, TRAIN_NN, 1], dim=1))


def bench_train(ctx):
    import torch
    from lpdnet_b200 import ops, optim, synth
    from lpdnet_b200 import train_pointnetvlad as TP
    from lpdnet_b200.loss import pointnetvlad_loss as PL
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    args, dist = ctx.args, ctx.dist
    ops.set_precision(args.precision)
    torch.manual_seed(1234)
    model = PointNetVlad(num_points=NPTS, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))          # identical replicas on every rank
    model = model.to(ctx.device).train()
    # lr 1e-7: with the reference's default 1e-3 (or even 1e-5) Adam drives the hinge loss of these 4 synthetic tuple batches
    # to exactly 0 within ~10 steps and every gradient becomes zero; the kernels do the same work either way (nothing is
    # data-dependent), but a tiny lr keeps the loss and gradients non-trivial for the whole timed region.
    opt = optim.Adam(model.parameters(), lr=1e-7)
    n_rot = 4
    host = [tuple(t.pin_memory() for t in synth_tuples(4321 + 131 * ctx.rank + i)) for i in range(n_rot)]
    dev_in = [tuple(t.to(ctx.device) for t in h) for h in host]
    state = {}

    def step(i):
        state["loss"] = TP.train_step(model, opt, *dev_in[i % n_rot], margin_1=0.5, margin_2=0.2)

    for i in range(ctx.warmup):
        step(i)
    ctx.barrier()
    ops.reset_launch_count()
    with ClockSampler(ctx.local) as clk:
        t_ms, per_rank, per_step = ctx.timed_steps(step)
    launches = ops.launch_count()
    value = ctx.world * TRAIN_CLOUDS * ctx.steps / (t_ms * 1e-3)
    last_loss = float(state["loss"])
    ranks = ctx.rank_report(per_step, clk.result)

    # ---- end to end: pinned host tuples -> H2D -> step -> loss value back on the host, every step ----
    def e2e():
        for i in range(ctx.steps):
            state["loss_host"] = float(TP.train_step(model, opt, *host[i % n_rot], margin_1=0.5, margin_2=0.2))
    te = ctx.timed_region(e2e)
    e2e_value = ctx.world * TRAIN_CLOUDS * ctx.steps / (te * 1e-3)

    # ---- N > 1: the all-reduced gradient equals the sum of the per-rank gradients ----
    selfcheck = None
    if ctx.world > 1:
        opt.keep_local = True                                         # keep this rank's own gradient next to the reduced one
        opt.zero_grad()
        o = TP.run_model(model, *dev_in[0])
        loss = PL.quadruplet_loss(*o, 0.5, 0.2, use_min=True, lazy=True, ignore_zero_loss=False)
        loss.backward()
        reduced = opt.reduce_gradients()                              # what step() does before lpd_adam: NCCL sum over the ranks
        local = opt.local_gradients()                                 # this rank's own contribution, kept aside by the optimizer
        parts = [torch.empty_like(local) for _ in range(ctx.world)]
        dist.all_gather(parts, local)
        total = parts[0].double()
        for p_ in parts[1:]:
            total += p_.double()
        scale = float(total.abs().max())
        diff = float((reduced.double() - total).abs().max())
        selfcheck = {"allreduced_gradient_equals_sum_of_rank_gradients": bool(diff <= 1e-5 * scale),
                     "max_abs_diff": diff, "gradient_max_abs": scale, "elements": int(local.numel()),
                     "ranks_differ": bool(float((parts[0] - parts[-1]).abs().max()) > 0.0)}
        opt.keep_local = False
        opt.zero_grad()

    # per-kernel device time of a step: every rank runs these steps (the all-reduce is a collective), rank 0 records
    tot, cnt = ctx.profile_step(step, 2, record=ctx.rank == 0)
    res = None
    if ctx.rank == 0:
        step_ms = sum(tot.values())
        breakdown = {k_: {"ms_per_step": round(v, 4), "share": round(v / step_ms, 4), "launches": cnt[k_]}
                     for k_, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]}
        top = max(tot, key=tot.get)
        nparams = sum(p.numel() for p in model.parameters())
        res = {"metric": TRAIN_METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": ctx.steps, "warmup": ctx.warmup,
               "ms_per_step": t_ms / ctx.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic",
               "config": {"workload": "C3: LPD-Net training step (train-mode forward, lazy quadruplet loss m1=0.5 m2=0.2, backward, "
                                      "gradient all-reduce, fused Adam), 2 tuples = 44 x 4096-pt submaps per GPU per step",
                          "precision": "fp32" if args.precision == "fp32" else "tf32 tensor-core GEMMs (the training path has no f16 form)",
                          "submaps_per_gpu_per_step": TRAIN_CLOUDS, "tuples_per_gpu": TRAIN_BQ, "points": NPTS,
                          "sharding": f"whole tuples per GPU (dp{ctx.world}), per-rank BatchNorm statistics; gradients: {opt.describe_reduction()} "
                                      f"({4 * nparams / 1e6:.1f} MB fp32 per step)",
                          "l2": "256 MiB memset between timed steps (untimed); 4 rotating tuple batches", "last_loss": last_loss},
               "clocks": clk.result, "gpu_launches": launches,
               "per_rank_ms_per_step": [round(v / ctx.steps, 4) for v in per_rank], "ranks": ranks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": TRAIN_CLOUDS * NPTS * 3 * 4, "d2h_bytes_per_step": 4,
                       "api": "lpdnet_b200.train_pointnetvlad.train_step (pinned host tuples -> loss value on the host)",
                       "last_loss": state.get("loss_host")},
               "roofline": roofline_of(top, tot[top], cnt[top], step_ms, ctx.peaks, TRAIN_CLOUDS),
               "kernel_breakdown": breakdown}
        if selfcheck is not None:
            res["selfcheck"] = selfcheck
    del model, opt
    torch.cuda.empty_cache()
    return res


# ======================================================================================================================
# C4: retrieval
# ======================================================================================================================
def bench_retrieval(ctx):
    import numpy as np
    import torch
    from lpdnet_b200 import evaluate, ops, parallel, synth
    DB, Q, SETS = synth.descriptor_database()
    R = len(DB)
    truth = evaluate.prepare_truth(SETS)                              # ground-truth lists -> CSR, once per evaluation set
    DBd = [torch.from_numpy(d).to(ctx.device) for d in DB]
    Qd = [torch.from_numpy(q).to(ctx.device) for q in Q]
    searches = sum(len(Q[n]) for n in range(R)) * (R - 1)             # one 25-NN search per (query, other run) = 66,792
    state = {}

    def step(i):
        state["res"] = evaluate.recall_all_pairs(DBd, Qd, SETS, truth)   # runs sharded over the ranks, counters all-reduced

    for i in range(ctx.warmup):
        step(i)
    ops.reset_launch_count()
    with ClockSampler(ctx.local) as clk:
        t_ms, per_rank, _ = ctx.timed_steps(step)
    launches = ops.launch_count()
    value = searches * ctx.steps / (t_ms * 1e-3)                      # fixed total work: strong scaling

    # ---- end to end: pinned host descriptors -> H2D -> search + counting -> recall tables on the host ----
    DBh = [torch.from_numpy(d).pin_memory() for d in DB]
    Qh = [torch.from_numpy(q).pin_memory() for q in Q]

    e2e_ms = []

    def e2e(n=None):
        for _ in range(n or ctx.steps):
            t0 = time.perf_counter()
            state["res_host"] = evaluate.recall_all_pairs([d.to(ctx.device, non_blocking=True) for d in DBh],
                                                          [q.to(ctx.device, non_blocking=True) for q in Qh], SETS, truth)
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
    e2e(2)                                                            # warm the host-buffer path
    e2e_ms.clear()
    te = ctx.timed_region(e2e)
    e2e_value = searches * ctx.steps / (te * 1e-3)

    # ---- one big database: 21,988 rows sharded over the ranks, all 3,036 queries, NCCL all-gather + lpd_topk_merge ----
    db_all, q_all = torch.cat(DBd, 0), torch.cat(Qd, 0)
    lo, hi = parallel.shard_range(db_all.shape[0], ctx.world, ctx.rank)
    shard = db_all[lo:hi].contiguous()

    def big(i):
        state["big"] = parallel.sharded_retrieval_topk(shard, q_all, 25, lo)
    for i in range(3):
        big(i)
    tb_ms, _, _ = ctx.timed_steps(big)
    big_qps = q_all.shape[0] * ctx.steps / (tb_ms * 1e-3)
    selfcheck = None
    if ctx.world > 1:
        ref_idx, ref_d = ops.retrieval_topk(db_all, q_all, 25)        # unsharded fp64 brute force on every rank
        idx, dst = state["big"]
        selfcheck = {"sharded_top25_equals_unsharded_search": bool(torch.equal(idx, ref_idx) and torch.equal(dst, ref_d)),
                     "queries": int(q_all.shape[0]), "database_rows": int(db_all.shape[0]), "shards": ctx.world}
        flag = torch.tensor([int(selfcheck["sharded_top25_equals_unsharded_search"])], device=ctx.device)
        ctx.dist.all_reduce(flag, op=ctx.dist.ReduceOp.MIN)
        selfcheck["on_every_rank"] = bool(int(flag))

    tot, cnt = ctx.profile_step(step, 2, record=ctx.rank == 0)
    res = None
    if ctx.rank == 0:
        g = None
        try:
            g = np.load(ROOT / "tests" / "golden" / "recall.npz", allow_pickle=False)
        except OSError:
            pass
        r = state["res"]
        off = ~np.eye(R, dtype=bool)
        recall1 = float(np.mean(r["recall"][off][:, 0]))
        one_pct = float(np.mean(r["one_pct"][off]))
        same = None
        if g is not None:   # golden order: for m: for n != m  -> [n][m]
            want = np.stack([g["recall"][p] for p in range(len(g["recall"]))])
            got = np.stack([r["recall"][n, m] for m in range(R) for n in range(R) if m != n])
            same = bool(np.array_equal(got, want) and np.array_equal(
                np.array([r["one_pct"][n, m] for m in range(R) for n in range(R) if m != n]), g["one_percent"]))
        step_ms = sum(tot.values())
        top = max(tot, key=tot.get)
        res = {"metric": C4_METRIC, "value": value, "unit": "searches/s", "n_gpus": ctx.world, "steps": ctx.steps, "warmup": ctx.warmup,
               "ms_per_step": t_ms / ctx.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32x3 filter + f64 refine",
               "data": "synthetic",
               "config": {"workload": "C4: recall@1..25 / recall@1% over all 506 ordered run pairs (23 runs x 956 database rows, 132 queries per run, "
                                      "256-d) in one pass per step: stacked databases -> 3xTF32 distance GEMM (tcgen05) -> threshold filter -> "
                                      "fp64 re-rank -> on-device first-hit histogram",
                          "searches_per_step": searches, "sharding": f"database runs round-robin over {ctx.world} rank(s); integer counters all-reduced",
                          "recall_at_1_percent": recall1, "recall_at_1pct_percent": one_pct,
                          "identical_to_reference_kdtree_golden": same,
                          "l2": "256 MiB memset between timed steps (untimed)"},
               "clocks": clk.result, "gpu_launches": launches, "per_rank_ms_per_step": [round(v / ctx.steps, 4) for v in per_rank],
               "e2e": {"value": e2e_value, "unit": "searches/s", "h2d_bytes_per_step": int(sum(d.numel() for d in DBh) * 4 / ctx.world + sum(q.numel() for q in Qh) * 4),
                       "d2h_bytes_per_step": R * R * 27 * 4 + int(q_all.shape[0]) * R * 4,
                       "api": "lpdnet_b200.evaluate.recall_all_pairs (pinned host descriptors -> recall tables on the host)",
                       "host_ms_per_step_min_median_max": [round(min(e2e_ms), 3), round(statistics.median(e2e_ms), 3), round(max(e2e_ms), 3)]},
               "roofline": roofline_of(top, tot[top], cnt[top], step_ms, ctx.peaks, 1),
               "big_database": {"value": big_qps, "unit": "queries/s", "ms_per_step": tb_ms / ctx.steps, "queries": int(q_all.shape[0]),
                                "database_rows": int(db_all.shape[0]),
                                "algorithmic_tflops": 2.0 * q_all.shape[0] * db_all.shape[0] * 256 / (tb_ms / ctx.steps * 1e-3) / 1e12,
                                "path": "lpdnet_b200.parallel.sharded_retrieval_topk: rows sharded, local exact top-25, NCCL all-gather, lpd_topk_merge"},
               "kernel_breakdown": {k_: {"ms_per_step": round(v, 4), "launches": cnt[k_]} for k_, v in sorted(tot.items(), key=lambda kv: -kv[1])}}
        if selfcheck is not None:
            res["selfcheck"] = selfcheck
    return res


# ======================================================================================================================
# arms
# ======================================================================================================================
def cpu_baselines(which, threads):
    """bounded CPU samples of the reference for the requested workloads (rank 0, N = 1 only)"""
    out = {}
    if "c2" in which:
        v, dt, kind, _ = cpu_embed(2 * 16, threads)
        out["c2"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                     "sample": f"32 submaps (of the 64-submap batch) in 2 chunks of 16, {dt:.1f} s, torch {threads} threads"}
    if "c3" in which:
        v, dt, kind = cpu_train_step(1, threads)
        out["c3"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                     "sample": f"1 tuple (22 of the 44 submaps): train-mode forward + loss + backward + Adam in {dt:.1f} s"}
    if "c4" in which:
        v, dt, kind, n = cpu_recall(100)
        out["c4"] = {"value": v, "unit": "searches/s", "cores": 1, "kind": kind,
                     "sample": f"100 of the 506 run pairs ({n} KDTree.query(k=25) calls + 100 tree builds) in {dt:.1f} s, single-threaded as written"}
    if "c5" in which:
        v, dt, kind, _ = cpu_embed(1, threads, chunk=1, npts=C5_NPTS, k=C5_KNN)
        out["c5"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                     "sample": f"1 cloud of 16384 points, k=32 (each [N,N] temporary is 1 GiB) in {dt:.1f} s, scaled linearly"}
    return out


def run_ours(args):
    ctx = Ctx(args)
    which = ["c2", "c3", "c4", "c5"] if args.workload == "all" else [args.workload]
    fns = {"c2": lambda: bench_embed(ctx, "c2"), "c3": lambda: bench_train(ctx), "c4": lambda: bench_retrieval(ctx),
           "c5": lambda: bench_embed(ctx, "c5")}
    results = {}
    t0 = time.perf_counter()
    for w in which:
        results[w] = fns[w]()
        if ctx.rank == 0:
            print(f"[bench] {w} done at {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    if ctx.rank == 0:
        ctx.peaks["tf32_tflops"] = measure_tf32_peak(ctx.device)
        for r in results.values():      # grade the tf32 kernels against their own measured peak as well
            for rf in [r.get("roofline")] + list(r.get("rooflines", [])):
                if rf and rf.get("bound") == "tensor":
                    rf["frac_of_tf32_peak"] = rf["achieved"] / ctx.peaks["tf32_tflops"]
                    rf["tf32_peak"] = ctx.peaks["tf32_tflops"]
        if ctx.world == 1 and not args.no_cpu_baseline:
            base = cpu_baselines(which, os.cpu_count() or 1)
            for w, b in base.items():
                results[w]["cpu_baseline"] = b
        head = which[0]
        line = results[head]
        line["peaks"] = {**ctx.peaks, "tf32_how": "torch.matmul fp32 with allow_tf32, 8192^3, best of 10 (measured in this run)"}
        line["host"] = {"cores_per_rank": len(ctx.cores), "cpu_count": os.cpu_count()}
        if len(which) > 1:
            line["workloads"] = {w: results[w] for w in which[1:]}
        emit(line)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def run_reference(args):
    """the reference's own CPU implementation, all host threads.  Top level: C2, each step = the 64-submap batch in 4 chunks of
    16 (bounded if the box is slow); workloads c3 / c4 / c5 measured once each on bounded samples."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    which = ["c2", "c3", "c4", "c5"] if args.workload == "all" else [args.workload]
    line = None
    if "c2" in which:
        v1, t1, kind, model = cpu_embed(16, threads)                      # warm-up chunk (page-in, thread pools) + cost probe
        budget = 200.0
        total_steps = args.steps + min(args.warmup, 1)
        chunks = 4
        while chunks > 1 and chunks * t1 * total_steps > budget:
            chunks -= 1
        n = 16 * chunks
        for _ in range(min(args.warmup, 1)):
            cpu_embed(n, threads, model=model)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_embed(n, threads, model=model)
        dt = time.perf_counter() - t0
        v = n * args.steps / dt
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: LPD-Net eval embedding (featnet=lpdnet, kNN k=20 graph features + NetVLAD K=64 D=1024 -> 256), 64 x 4096-pt submaps per GPU per step",
                           "step": f"{n} submaps per step in chunks of 16 on the host CPU ({'the whole 64-submap batch' if n == 64 else 'bounded sample of the 64-submap batch'}); "
                                   f"1 warm-up step", "submaps_per_gpu_per_step": n, "points": NPTS, "k": KNN},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                 "sample": f"{n} submaps/step x {args.steps} steps, " + ("the unmodified reference modules (oracle/_ref), torch CPU"
                                                                                          if kind == "reference" else "numpy+C port of the reference's as-written forward")},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    rest = [w for w in which if w != "c2"] if line is not None else which[1:]
    others = cpu_baselines(rest if line is not None else which, threads)
    if line is None:
        w = which[0]
        b = others[w]
        line = {"impl": "reference", "metric": {"c3": TRAIN_METRIC, "c4": C4_METRIC, "c5": C5_METRIC}[w], "value": b["value"], "unit": b["unit"],
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": w, "step": b["sample"]}, "cpu_baseline": b,
                "e2e": {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        line["workloads"] = {w: {"impl": "reference", "value": b["value"], "unit": b["unit"], "cpu_baseline": b} for w, b in others.items()}
    emit(line)


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
    ap.add_argument("--precision", default="f16", choices=["f16", "tf32", "fp32"],
                    help="f16: fp16 activations / operands downstream of the kNN on tcgen05 kind::f16, fp32 accumulation (eval path); "
                         "tf32: dense layers on tcgen05 kind::tf32; fp32: strict mode.  The descriptor error vs the reference is measured "
                         "and reported in config.")
    ap.add_argument("--workload", default="all", choices=["all", "c2", "c3", "c4", "c5"],
                    help="all (default): C2 as the top-level line (the configuration BASELINE.json's metric is quoted on) with C3, C4, C5 "
                         "under `workloads`; or one workload as the top-level line")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-CPU reference samples (development)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
