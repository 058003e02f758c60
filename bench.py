#!/usr/bin/env python
"""bench.py -- grasps/sec of the PointNet grasp-quality training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic grasp clouds: forward
(PointNetCls, train mode) + nll_loss + backward + the NCCL gradient all-reduce (N>1) + Adam step,
exactly the body of main_1v.py:72-76, at BASELINE config 2: B=512 clouds/GPU x N=1024 points, k=2.
Weak scaling: every rank processes its own 512-cloud batch.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is produced.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grasps_per_sec_fwd_bwd"
UNIT = "grasps/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def fwd_flops_per_grasp(N, k):
    """Dense algorithmic flops of one forward."""
    return N * 557842 + 2626048 + 512 * k


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(B_sample, N, k, steps, warmup, max_seconds=None):
    """Time the oracle's torch port (the reference's op sequence on ATen/MKLDNN, all host threads):
    forward + nll_loss + backward + Adam, on a bounded sample of the workload (B_sample clouds)."""
    from oracle import pointnet_torch_port as PT
    from oracle import weights as W
    import torch.nn.functional as F
    st = W.make_state(0, k=k)
    sd = PT.to_torch_state(st, torch.float32, requires_grad=True)
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=0.005)
    x = torch.tensor(W.make_clouds(1234, B_sample, N, "box"))
    y = torch.tensor(W.make_labels(4321, B_sample, k))
    times = []

    def one():
        opt.zero_grad(set_to_none=True)
        logp, _ = PT.pointnetcls_forward(sd, x, training=True)
        loss = F.nll_loss(logp, y)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        one()
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if max_seconds is not None and time.perf_counter() - t_all > max_seconds:
            break
    ms = 1e3 * sum(times) / len(times)
    return {"value": B_sample / (ms / 1e3), "ms_per_step": ms, "steps_done": len(times),
            "cores": torch.get_num_threads()}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="clouds per GPU (BASELINE config 2: 512)")
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--simt", action="store_true", help="force the fp32 CUDA-core kernels")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=128)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, N, k = args.batch, args.points, args.classes
    workload = "PointNetCls k=%d train step (fwd+nll+bwd+allreduce+Adam), B=%d clouds/GPU x N=%d pts" % (k, B, N)
    config = {"workload": workload, "global_batch": B * world, "points": N, "classes": k,
              "parallelism": "dp%d" % world, "cuda_graph": None, "l2": "per-step working set ~1.2 GB of saved activations >> 126 MB L2; 8 rotating input batches"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        warm = max(1, min(args.warmup, 2))
        steps = max(1, args.steps)
        r = cpu_reference_arm(args.cpu_sample_batch, N, k, steps, warm, max_seconds=150.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": r["steps_done"], "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": "oracle torch port (reference op sequence on ATen CPU), %d-cloud batches of %d points, "
                                           "fwd+nll+bwd+Adam per step" % (args.cpu_sample_batch, N)},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch.distributed as dist
    from oracle import weights as W   # deterministic synthetic weights / clouds (data generator only)
    from pointnetgpd_b200 import _abi as A
    from pointnetgpd_b200.ddp import FlatGradAllReduce
    from pointnetgpd_b200.model.pointnet import PointNetCls
    import torch.nn.functional as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (the fused path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = A.load()

    st = W.make_state(0, k=k)
    model = PointNetCls(num_points=N, input_chann=3, k=k)
    model.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()})
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=0.005, fused=True, capturable=True)
    sync = FlatGradAllReduce(list(model.parameters()), world)
    flags_extra = A.F_SIMT if args.simt else 0
    if flags_extra:
        from pointnetgpd_b200.functional import run_module
        fwd = lambda xx: run_module(model, A.PGPD_CLS, xx, k=k, flags_extra=flags_extra)
    else:
        fwd = model

    NBUF = 8
    xs_host = [torch.tensor(W.make_clouds(1234 + rank * 100 + i, B, N, "box")).pin_memory() for i in range(NBUF)]
    ys_host = [torch.tensor(W.make_labels(4321 + rank * 100 + i, B, k)).pin_memory() for i in range(NBUF)]
    xs_dev = [t.to(dev) for t in xs_host]
    ys_dev = [t.to(dev) for t in ys_host]

    def eager_step(x, y):
        opt.zero_grad(set_to_none=True)
        logp, _ = fwd(x)
        loss = F.nll_loss(logp, y)
        loss.backward()
        sync.all_reduce()
        opt.step()
        return loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The step is captured once into a CUDA graph and replayed (pointnetgpd_b200.graph.GraphedTrainStep); the dominant
    # kernel's event pair is recorded INSIDE the graph (external event nodes), so its duration is still measured live.
    use_graph = not args.no_graph and not args.simt
    graphed = None
    lib.pgpd_profile_enable(0)
    if use_graph:
        try:
            from pointnetgpd_b200.graph import GraphedTrainStep
            graphed = GraphedTrainStep(model, opt, xs_dev[0], ys_dev[0], grad_sync=sync if world > 1 else None, warmup=3,
                                       before_capture=lambda: lib.pgpd_profile_enable(2))
        except Exception as e:           # e.g. a collective that cannot be captured: fall back to eager launches
            sys.stderr.write("CUDA-graph capture failed (%s: %s); running eagerly\n" % (type(e).__name__, e))
            graphed = None
            lib.pgpd_profile_enable(0)
    step = graphed.step if graphed is not None else eager_step

    for i in range(args.warmup):
        step(xs_dev[i % NBUF], ys_dev[i % NBUF])
    barrier()

    # ---- device-timed region: inputs resident in HBM -------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if graphed is None:
        lib.pgpd_profile_enable(1)
    n0 = lib.pgpd_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(xs_dev[i % NBUF], ys_dev[i % NBUF])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.pgpd_launch_count() - n0)
    import ctypes
    nl, tot = ctypes.c_int(0), ctypes.c_float(0.0)
    lib.pgpd_profile_read(ctypes.byref(nl), ctypes.byref(tot))      # graph mode: the event nodes of the LAST replayed step
    lib.pgpd_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    if graphed is not None:
        # replays launch the captured kernels without passing through the library's host code: count them from one eager step
        n1 = lib.pgpd_launch_count()
        eager_step(xs_dev[0], ys_dev[0])
        torch.cuda.synchronize(dev)
        launches = int(lib.pgpd_launch_count() - n1) * args.steps

    # ---- end-to-end region: pinned host inputs -> H2D -> step -> D2H of the loss ------------------
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        if graphed is not None:
            loss = graphed.step(xs_host[i % NBUF], ys_host[i % NBUF])      # pinned host -> static device buffers -> replay
        else:
            x = xs_host[i % NBUF].to(dev, non_blocking=True)
            y = ys_host[i % NBUF].to(dev, non_blocking=True)
            loss = eager_step(x, y)
        _ = loss.item()                     # D2H read of the step's result
    barrier()
    e2e_ms_total = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([ms_total, e2e_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = B * world / (ms_per_step / 1e3)
    e2e_value = B * world / (e2e_ms_total / args.steps / 1e3)

    config["cuda_graph"] = graphed is not None
    if rank == 0:
        peaks, peak_src = load_peaks()
        M = B * N
        k3_flops = 2.0 * 128 * 1024 * M                       # layer-3 GEMM of one tower forward, algorithmic
        k3_ms = (tot.value / nl.value) if nl.value else None
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        traffic = None          # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        tpath = os.path.join(ROOT, "profiles", "r1_l3_traffic.json")
        if os.path.exists(tpath) and (B, N) == (512, 1024) and not args.simt:
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "tower layer-3 GEMM (128->1024) + max-pool epilogue, one launch per tower forward",
                    "achieved": (k3_flops / (k3_ms * 1e-3) / 1e12) if k3_ms else None, "peak": peak, "unit": "TFLOP/s",
                    "frac": ((k3_flops / (k3_ms * 1e-3) / 1e12) / peak) if k3_ms else None,
                    "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "kernel_ms": k3_ms, "launches_timed": nl.value, "traffic": traffic,
                    "impl": "tcgen05" if (lib.pgpd_has_tensor_core_path() and not args.simt) else "cuda-core fp32",
                    "step_algorithmic_tflops": 3.0 * fwd_flops_per_grasp(N, k) * B / (ms_per_step * 1e-3) / 1e12,
                    "timing": ("event pairs recorded as nodes of the replayed CUDA graph (last timed step)" if graphed is not None
                               else "event pairs around every launch in the timed region")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * N * 4 + B * 8, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms_total / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_arm(args.cpu_sample_batch, N, k, steps=3, warmup=1, max_seconds=25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": "oracle torch port on the host CPU: %d timed fwd+nll+bwd+Adam steps on a %d-cloud x %d-point "
                                              "batch (1 warm-up)" % (r["steps_done"], args.cpu_sample_batch, N)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the line is out; never let communicator teardown hang the job
        threading.Timer(30.0, lambda: os._exit(0)).start()
        try:
            del graphed
            torch.cuda.synchronize(dev)
            dist.barrier()
            dist.destroy_process_group()
        finally:
            sys.stdout.flush()
            os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
