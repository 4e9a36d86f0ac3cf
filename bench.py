#!/usr/bin/env python
"""bench.py -- samples/sec of `forward_kld` on the flagship neural-spline stack (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5          # ours, one B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # data parallel: batch sharded, one all-reduce
    python bench.py --impl reference ...                     # CPU arm: the oracle port on the host cores

Workload (configs[1] of BASELINE.json): 32 x [AutoregressiveRationalQuadraticSpline(64, 2
blocks, 256 hidden, 8 bins, tail 3) + LULinearPermute(64)], DiagGaussian(64) base, batch 65 536 per GPU,
fp32 in/out, synthetic inputs x = 1.5 * randn, random-init weights moved off identity-init
(sigma 0.03 on the conditioners, 0.01 on the LU factors) so that splines/tails are non-degenerate.
One "step" = one full `forward_kld` pass over one batch (all 64 layers + base density + mean).

Timed region: W warm-up steps, then exactly K steps between barrier+synchronize, CUDA events on the
launching stream, max over ranks.  Inputs rotate through NBUF distinct device batches whose total size
exceeds L2 (so no step re-reads its input from L2); the packed weights (85 MB) cycle through L2 as they
do in real use.  `e2e` is the same pass through the C-ABI host entry point (`nfb_flow_forward_kld_host`):
pinned host batch -> H2D -> kernels -> D2H of the scalar loss, every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200")]

D, LAYERS, HIDDEN, BLOCKS, BINS, TAIL = 64, 32, 256, 2, 8, 3.0
BATCH = 65536
KIND = os.environ.get("NFB_BENCH_KIND", "ar")  # "ar" (BASELINE config 2) or "coupled"
# algorithmic work per sample per [spline block + LU]: GEMM flops 2*sum(in*out)
FLOPS_PER_SAMPLE_LAYER = {"ar": 1_327_104, "coupled": 933_888}
MIN_BYTES_PER_SAMPLE_LAYER = 520  # z in + z out + log_q r/w
METRIC = "samples/sec forward_kld, 32-layer RQ-NSF d=64 batch=65536"


def workload_name(batch=BATCH):
    return (f"{'Autoregressive' if KIND == 'ar' else 'Coupled'} RQ-NSF d={D}, {LAYERS} x "
            f"[spline block(2 blocks, hidden {HIDDEN}, {BINS} bins) + LULinearPermute], "
            f"batch {batch}/GPU, forward_kld (BASELINE.json configs[1])")


def build_model(kind=KIND, layers=LAYERS, seed=0):
    import torch
    import normflows as nf
    torch.manual_seed(seed)
    fl = []
    for i in range(layers):
        if kind == "ar":
            fl.append(nf.flows.AutoregressiveRationalQuadraticSpline(D, BLOCKS, HIDDEN, num_bins=BINS, tail_bound=TAIL))
        else:
            fl.append(nf.flows.CoupledRationalQuadraticSpline(D, BLOCKS, HIDDEN, num_bins=BINS, tail_bound=TAIL,
                                                             reverse_mask=bool(i % 2)))
        fl.append(nf.flows.LULinearPermute(D))
    model = nf.NormalizingFlow(nf.distributions.DiagGaussian(D, trainable=False), fl)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.add_((0.01 if ".linear." in name else 0.03) * torch.randn(p.shape, generator=g))
    return model


def oracle_spec(kind=KIND, layers=LAYERS):
    t = "AutoregressiveRationalQuadraticSpline" if kind == "ar" else "CoupledRationalQuadraticSpline"
    return {"kind": "NormalizingFlow", "q0": {"shape": [D]},
            "flows": [{"type": t, "num_bins": BINS, "tail_bound": TAIL}, {"type": "LULinearPermute"}] * layers}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [s.strip() for s in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def _cpu_worker(args):
    kind, rows, reps, blas_threads, seed = args
    import numpy as np
    from threadpoolctl import threadpool_limits
    from oracle import nf_oracle as O
    import torch
    torch.set_num_threads(1)
    model = build_model(kind)
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    spec = oracle_spec(kind)
    x = (np.random.default_rng(seed).normal(size=(rows, D)) * 1.5).astype(np.float32)
    with threadpool_limits(limits=blas_threads):
        O.forward_kld(spec, sd, x[:64])  # warm
        t0 = time.time()
        for _ in range(reps):
            kld = O.forward_kld(spec, sd, x)
        dt = time.time() - t0
    return dt, float(kld)


def cpu_baseline(seconds_target=15.0, kind=KIND, rows=512, reps=2):
    """The oracle port (numpy restatement of the reference algorithm) on ALL of this host's cores, on a
    bounded sample of the SAME workload: the batch is data-parallel, so `cores // 8` worker processes
    each push `rows` samples through the full 32-layer stack with 8 BLAS threads."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per = 8 if cores >= 8 else cores
    workers = max(1, cores // per)
    ctx = mp.get_context("spawn")
    t0 = time.time()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(kind, rows, reps, per, 1234 + i) for i in range(workers)])
    wall = time.time() - t0
    dt = max(r[0] for r in res)  # slowest worker's compute time for reps passes
    value = workers * rows * reps / dt
    return {"value": value, "unit": "samples/s", "cores": workers * per, "kind": "port",
            "sample": f"oracle/nf_oracle.py (numpy fp32) forward_kld, {workers} processes x {per} BLAS threads, "
                      f"{rows} rows x {LAYERS} layers x {reps} passes each ({wall:.1f} s wall incl. start-up); "
                      f"kld={res[0][1]:.4f}"}, dt / reps


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU (oracle port; the Python reference
    package itself does not travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    # one step = one pass of a bounded sample (rows per worker) through the full 32-layer stack on all host cores;
    # the sample shrinks with the step count so that warm-up + K steps stay within a couple of minutes
    rows = 512 if steps <= 4 else (256 if steps <= 12 else 96)
    base, dt = cpu_baseline(rows=rows, reps=steps)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "sample": f"{rows} rows per worker process per step"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="rows per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import normflows as nf
    from normflows.parallel import forward_kld_dp
    torch.set_grad_enabled(False)  # the metric is the forward pass (timed under torch.no_grad())

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"launched {world} ranks for --gpus {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # NCCL sets channels up lazily over the first collectives (hundreds of ms when the GPU is busy):
        # keep that out of the warm-up/timed steps
        t = torch.ones(2, dtype=torch.float64, device=dev)
        for _ in range(20):
            dist.all_reduce(t)
        torch.cuda.synchronize()
        dist.barrier()
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    B = args.batch

    model = build_model().to(dev)
    nbuf = max(2, (160 << 20) // (B * D * 4) + 1)  # rotating inputs: > 126 MB L2 in total
    g = torch.Generator().manual_seed(1234 + rank)
    xs_host = [(torch.randn(B, D, generator=g) * 1.5).pin_memory() for _ in range(2)]
    xs = [(torch.randn(B, D, generator=g) * 1.5).to(dev) for _ in range(nbuf)]

    dp_mode = os.environ.get("NFB_BENCH_DP", "async")  # async | sync | none(debug: no collective)

    def step(i):
        if dp_mode == "none":
            return model.forward_kld(xs[i % nbuf])
        return forward_kld_dp(model, xs[i % nbuf], async_op=(dp_mode == "async"))

    def value_of(l):
        return l.result() if hasattr(l, "result") else l

    for i in range(warmup):
        loss = value_of(step(i))
    stack = model._stack()
    launches_per_step = stack.launch_count()
    fused = stack.fused_layers()
    assert len(fused) == 2 * LAYERS, "flagship stack must run on the fused tcgen05 kernel"
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    t_host0 = time.perf_counter()
    for i in range(steps):
        loss = step(i)
    loss = value_of(loss)  # stream-waits for the last collective; every step's loss was reduced on device
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps  # CPU time to enqueue one step
    ev1.record()
    torch.cuda.synchronize()
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        dist.barrier()
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t)
    loss_val = float(loss)

    # ---- e2e through the host-buffer C-ABI entry point (H2D + kernels + D2H each step) ----
    for i in range(3):
        model.forward_kld_host(xs_host[i % 2], dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_loss = model.forward_kld_host(xs_host[i % 2], dev)  # synchronous: returns the host float
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (fused spline block), timed live with CUDA events ----
    roof = None
    if rank == 0:
        import ctypes as C
        from normflows import _lib as L
        # The dominant kernel is the persistent whole-stack launch (all 32 [LU + spline block] pairs, one
        # kernel): time it alone, back to back, inputs rotating as above.
        h = stack._h
        n_l = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outs = [torch.empty_like(xs[0]) for _ in range(2)]
        ld = torch.zeros(B, device=dev)
        for i in range(2):
            L.check(L.lib().nfb_flow_transform(h, L.NFB_INVERSE, L.ptr(xs[i % nbuf]), L.ptr(outs[i % 2]), L.ptr(ld),
                                               B, L.stream_ptr()))
        torch.cuda.synchronize()
        e0.record()
        for i in range(n_l):
            L.check(L.lib().nfb_flow_transform(h, L.NFB_INVERSE, L.ptr(xs[i % nbuf]), L.ptr(outs[i % 2]), L.ptr(ld),
                                               B, L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / n_l  # includes a 65 K-element fill and a 2 KB memset (~4 us)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        flops = FLOPS_PER_SAMPLE_LAYER[KIND] * B * LAYERS
        achieved = flops / (k_ms * 1e-3) / 1e12
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        roof = {"kernel": "nfb::fused_rqs_kernel, whole stack in one persistent launch: 32 x (LULinearPermute + MADE "
                          "conditioner + RQ spline + log-det), (layer, tile) work units",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel_ms": k_ms,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
                "note": "algorithmic = dense fp32-equivalent GEMM flops of the reference (1.327 MFLOP/sample/layer x 32 "
                        "layers, SURVEY 8d).  The kernel runs every product as 3 bf16 tensor-core passes (split "
                        "precision, needed for the rtol 1e-4 bar) so frac <= 1/3 by construction, and skips the "
                        "all-zero blocks of the MADE masks (~31 % of the dense MMA work); ncu: tensor pipe 44 % active"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B * steps / (elapsed_ms * 1e-3)
    e2e_value = world * B * steps / e2e_s
    line = {"metric": METRIC, "value": value,
            "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B),
                       "global_batch": world * B, "parallelism": f"dp{world}", "dp_collective": dp_mode,
                       "l2_policy": f"{nbuf} rotating input batches ({nbuf * B * D * 4 >> 20} MiB > L2)",
                       "loss": loss_val},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * D * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s / steps * 1e3, "loss": e2e_loss,
                    "api": "nfb_flow_forward_kld_host (pinned host batch)"},
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "gpu_launches": launches_per_step * steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roof}
    if not args.no_cpu_baseline and world == 1:  # reported baseline: rank 0 at N = 1 only
        line["cpu_baseline"], _ = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
