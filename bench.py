"""Benchmark of the ProteinReDiff denoiser hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

metric   denoiser steps/sec at N=512 tokens, pair_dim 64, batch 8 (paper dims 512/64/4 blocks):
         one "step" = one evaluation of the denoising network on a batch of 8 complexes of 512
         tokens plus the on-device DDPM update, i.e. one iteration of the reference's sampling
         loop (model.py:403-420).
value    whole-job steps/sec, inputs resident in HBM, the captured CUDA graph replayed K times,
         CUDA events on the replay stream, max over ranks (weak scaling: every rank runs its own
         batch of 8, sample-parallel -- no data-path collective).
e2e      the same metric through the public API (model.sample_step) with HOST buffers: pinned
         host -> device copies of (z, seq_t, mask, t) and device -> host reads of
         (noise_pred, seq_pred) inside the timed region (sample_step replays its own cached graph).
roofline the dominant kernel (triangle-attention core), timed alone inside this script.
cpu_baseline / --impl reference
         the CPU oracle (a port of the reference's PyTorch forward, oracle/denoiser_ref.py) on the
         host cores, on a bounded sample: one complex (B=1) of the same N=512 workload.
gpu_eager_baseline
         the same port run as eager fp32 PyTorch (stock ATen / cuBLAS kernels) on the same GPU at the full
         batch of 8 (SURVEY §8d's second baseline; N=1 only, --no-gpu-eager skips it).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from protein_redesign_b200 import synthetic as syn  # noqa: E402

N_TOKENS, N_ATOMS, BATCH = 512, 32, 8
METRIC = "denoiser steps/sec (N=512 tokens, pair_dim 64, batch 8)"
UNIT = "steps/s"
WORKLOAD = "paper config single_dim 512/pair_dim 64/4 blocks, 512-token complex (32 ligand atoms + 480 residues), batch 8 per GPU"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_ram_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_reference_step_time(steps: int, warmup: int, budget_s: float = 150.0, rows: int = 1):
    """Time the CPU oracle (port of the reference forward) on `rows` complexes of the workload (1 = bounded sample,
    8 = the full batch: ~45 GB of activations)."""
    from oracle import denoiser_ref as ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = syn.PAPER
    sd = syn.make_state_dict(cfg, 0)
    batch = syn.make_batch(cfg, [(N_ATOMS, N_TOKENS - N_ATOMS)] * rows, seed=0)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, 0)
    torch.manual_seed(0)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if i >= warmup and sum(times) > budget_s:
                break
    return sum(times) / len(times), len(times), cores


def gpu_eager_step_time(dev, B: int, steps: int = 3, warmup: int = 1):
    """Second baseline: the reference's algorithm as eager fp32 PyTorch (the oracle port, stock ATen /
    cuBLAS kernels, torch's default matmul precision) on the SAME B200, timed with CUDA events.  Not the product path."""
    from oracle import denoiser_ref as ref

    cfg = syn.PAPER
    sd = {k: v.to(dev) for k, v in syn.make_state_dict(cfg, 0).items()}
    batch = syn.make_batch(cfg, [(N_ATOMS, N_TOKENS - N_ATOMS)] * B, seed=0)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, 0)
    torch.manual_seed(0)
    pb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in ref.prepare_batch(batch, cfg.mask_prob).items()}
    z, seq_t, mask, t = z.to(dev), seq_t.to(dev), mask.to(dev), t.to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.inference_mode():
        for _ in range(warmup):
            ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
        e1.record()
        torch.cuda.synchronize(dev)
    peak = torch.cuda.max_memory_allocated(dev)
    return e0.elapsed_time(e1) / steps * 1e-3, peak


def gpu_eager_baseline(dev):
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    try:
        try:
            t_step, peak = gpu_eager_step_time(dev, BATCH)
            rec = {"value": 1.0 / t_step, "sample": f"3 timed steps (1 warm-up) at B={BATCH}, N={N_TOKENS}"}
        except torch.OutOfMemoryError:
            torch.cuda.empty_cache()
            t_step, peak = gpu_eager_step_time(dev, 1)
            rec = {"value": 1.0 / (BATCH * t_step),
                   "sample": f"3 timed steps (1 warm-up) at B=1, N={N_TOKENS} (B={BATCH} does not fit); value = 1 / ({BATCH} * {t_step * 1e3:.1f} ms)"}
        rec.update({"unit": UNIT, "kind": "port", "device": torch.cuda.get_device_name(dev), "dtype": "f32 (torch defaults)",
                    "peak_mem_gb": round(peak / 1e9, 1),
                    "what": "eager PyTorch port of the reference forward (oracle/denoiser_ref.py) on the same GPU"})
        return rec
    except Exception as e:  # a baseline must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    finally:
        torch.cuda.empty_cache()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1)
    full = host_ram_gb() >= 80.0 and not args.reference_b1
    if full:
        # the real configuration: the whole batch of 8 on the host cores, one warm-up + at most two timed steps
        # (a step takes ~50 s on 16 cores) so that the run ends within minutes
        t_b8, n, cores = cpu_reference_step_time(min(args.steps, 2), warm, budget_s=60.0, rows=BATCH)
        t_step = t_b8 / BATCH
        sample = (f"{n} timed step(s) (after {warm} warm-up) of the CPU oracle on the FULL batch (B=8) of the N=512 workload, "
                  f"{t_b8:.1f} s per step")
    else:
        t_step, n, cores = cpu_reference_step_time(args.steps, warm)
        sample = (f"{n} timed steps (after {warm} warm-up) of the CPU oracle on ONE complex (B=1) of the N=512 workload "
                  f"(B=8 was NOT run on the CPU: needs ~45 GB); value = 1 / (8 * {t_step:.2f} s)")
    value = 1.0 / (BATCH * t_step)  # batch-of-8 steps per second
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": warm, "ms_per_step": BATCH * t_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _init_dist(dev):
    import torch.distributed as dist
    # NCCL prints its version banner to STDOUT when the communicator is created (NCCL_DEBUG=VERSION on some
    # boxes); the contract is ONE JSON line on stdout, so the banner goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class StepGraph:
    """One captured sampling step (denoiser + on-device DDPM update) on resident inputs, as ProteinReDiffModel.sample
    builds it (model.py: sample); replayable any number of times (the sampler state wraps after T steps)."""

    def __init__(self, model, cfg, host_batch, dev, seed, lib):
        from protein_redesign_b200 import ops
        torch.manual_seed(seed)
        self.batch = model.prepare_batch({k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()})
        self.host_inputs = syn.make_step_inputs(host_batch, cfg.num_steps, seed)
        z_h, seq_h, mask_h, _ = self.host_inputs
        B, N = mask_h.shape
        self.B, self.N = B, N
        ops.reserve_workspace(cfg, B, N, dev)
        self.z, self.seq_t, self.mask = z_h.to(dev), seq_h.to(dev), mask_h.to(dev)
        T = cfg.num_steps
        self.noise = torch.randn(T, B, N, 3, device=dev)
        ops.remove_mean(cfg, self.noise.view(-1, N, 3), self.mask)
        self.state = torch.tensor([T - 1, 0], dtype=torch.int32, device=dev)
        self.bufs = {"single": torch.empty(B, N, cfg.single_dim, device=dev), "pair": torch.empty(B, N, N, cfg.pair_dim, device=dev),
                     "opm_a": torch.empty(B, N, cfg.single_dim // 4, device=dev), "opm_b": torch.empty(B, N, cfg.single_dim // 4, device=dev),
                     "noise_pred": torch.empty(B, N, 3, device=dev), "seq_pred": torch.empty(B, N, 21, device=dev)}
        self.model, self.cfg, self.dev = model, cfg, dev
        with torch.inference_mode():
            model._static_embeddings(self.batch)
            self.one_step()
            torch.cuda.synchronize()
            c0 = lib.prd_launch_count()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                self.ws = ops.reserve_workspace(cfg, B, N, dev)  # the capture stream's scratch: kept alive with the graph
                with torch.cuda.graph(self.graph, stream=side):
                    self.one_step()
            self.launches_per_step = int(lib.prd_launch_count() - c0)
            torch.cuda.current_stream(dev).wait_stream(side)

    def one_step(self):
        from protein_redesign_b200 import ops
        eps, sp = self.model._denoise(self.batch, self.z, self.seq_t, self.mask, None, bufs=self.bufs, sampler_state=self.state)
        ops.sampler_update(self.cfg, eps, sp, self.noise, self.model._coef, self.z, self.seq_t, self.state)

    def time_replays(self, n, warmup, barrier=None, sampler=None):
        """ms for n replays (CUDA events on the current stream, synchronised on both sides)."""
        for _ in range(warmup):
            self.graph.replay()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            self.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        return e0.elapsed_time(e1)


def _make_model(cfg, dev, train=False):
    from protein_redesign_b200.model import ProteinReDiffModel
    model = ProteinReDiffModel(cfg)
    model.load_state_dict(syn.make_state_dict(cfg, 0), strict=True)
    model = model.to(dev)
    model = model.train() if train else model.eval()
    model.run_setup_schedule()
    model.setup_schedule = True
    return model


def run_ours(args):
    import torch.distributed as dist

    from protein_redesign_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        _init_dist(dev)
    lib = _lib.load()
    lib.prd_launch_count.restype = ctypes.c_longlong
    barrier = dist.barrier if world > 1 else None
    if args.workload == "train":
        return run_train(args, dev, rank, local_rank, world, lib)
    if args.workload != "sample":
        return run_config(args, dev, rank, lib)

    cfg = syn.PAPER
    model = _make_model(cfg, dev)
    host_batch = syn.make_batch(cfg, [(N_ATOMS, N_TOKENS - N_ATOMS)] * BATCH, seed=100 + rank)
    sg = StepGraph(model, cfg, host_batch, dev, rank, lib)
    B, N = sg.B, sg.N
    z_h, seq_h, mask_h, t_h = sg.host_inputs
    batch, mask, bufs = sg.batch, sg.mask, sg.bufs

    if args.profile_eager:
        with torch.inference_mode():
            for _ in range(args.warmup + args.steps):
                sg.one_step()
            torch.cuda.synchronize()
        print(json.dumps({"profile_eager": True, "steps": args.steps, "warmup": args.warmup}))
        return

    with torch.inference_mode():
        sampler = ClockSampler(local_rank) if rank == 0 else None
        ms = sg.time_replays(args.steps, args.warmup, barrier, sampler)
        clocks = sampler.stop() if rank == 0 else None
        launches_per_step = sg.launches_per_step

        # ---- sustained: the real use is 1000-step sampling -- >= 300 replays / >= 8 s with the clocks sampled throughout ----
        sustained = None
        if not args.no_sustained:
            n_sus = max(300, int(8.5e3 / max(ms / args.steps, 1e-3)))
            sampler2 = ClockSampler(local_rank) if rank == 0 else None
            ms_sus = sg.time_replays(n_sus, 0, barrier, sampler2)
            ck2 = sampler2.stop() if rank == 0 else None
            sustained = {"replays": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus, "clocks": ck2}

        # ---- ragged: the same shape with ~10 % padding per row (masked key tiles, padded rows everywhere) ----
        ragged = None
        if not args.no_ragged and rank == 0:
            sizes = [(N_ATOMS - 2 * (i % 3), N_TOKENS - N_ATOMS - 40 - 7 * i) for i in range(BATCH)]
            rb = syn.make_batch(cfg, sizes, seed=300, n_total=N_TOKENS)
            rg = StepGraph(model, cfg, rb, dev, 7, lib)
            ms_r = rg.time_replays(args.steps, args.warmup)
            ragged = {"ms_per_step": ms_r / args.steps, "value": 1e3 / (ms_r / args.steps), "unit": UNIT,
                      "valid_tokens_per_row": [a + r for a, r in sizes], "n_total": N_TOKENS}
            del rg

        # ---- end-to-end through the public API with host buffers ---------------------------------
        pin = lambda x: x.clone().pin_memory()
        hz, hs, hm, ht = pin(z_h), pin(seq_h), pin(mask_h), pin(t_h)
        out_n = torch.empty(B, N, 3).pin_memory()
        out_s = torch.empty(B, N, 21).pin_memory()
        h2d = sum(x.numel() * x.element_size() for x in (hz, hs, hm, ht))
        d2h = out_n.numel() * 4 + out_s.numel() * 4

        def e2e_step():
            dz, ds = hz.to(dev, non_blocking=True), hs.to(dev, non_blocking=True)
            dm, dt = hm.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
            n_, s_ = model.sample_step(batch, dz, ds, dm, dt)
            out_n.copy_(n_, non_blocking=True)
            out_s.copy_(s_, non_blocking=True)

        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        k_e2e = max(3, min(args.steps, 20))
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(k_e2e):
            e2e_step()
        e3.record()
        torch.cuda.synchronize()
        ms_e2e = e2.elapsed_time(e3) / k_e2e

        # ---- dominant kernel alone (roofline) ------------------------------------------------------
        roof = None
        if rank == 0 and hasattr(lib, "prd_profile_kernel"):
            roof = profile_dominant(lib, cfg, B, N, dev, mask, bufs["pair"], model)

    # ---- BASELINE config 3 for real: 64 samples of ONE 512-token complex, sharded rank::world, 8 per micro-batch,
    #      prepare_batch on the full batch before sharding, final NCCL all_gather (scripts/predict_batch_*.py:209-229) ----
    sp = None
    if not args.no_sample_parallel:
        sp = run_sample_parallel(model, dev, rank, world, t_steps=args.sp_steps)

    if world > 1:
        t_all = torch.tensor([ms, ms_e2e, sustained["seconds"] * 1e3 if sustained else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t_all[0]), float(t_all[1])
        if sustained:
            sustained["seconds"] = float(t_all[2]) * 1e-3
            sustained["ms_per_step"] = float(t_all[2]) / sustained["replays"]
    if rank == 0:
        ms_per_step = ms / args.steps
        value = world * 1e3 / ms_per_step
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            t_step, n, cores = cpu_reference_step_time(1, 1, budget_s=60.0)
            cpu = {"value": 1.0 / (BATCH * t_step), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n} timed step (after 1 warm-up) of the CPU oracle on ONE complex (B=1) of the N=512 workload "
                             f"(B=8 is timed by --impl reference); value = 1 / (8 * {t_step:.2f} s)"}
        eager = gpu_eager_baseline(dev) if world == 1 and not args.no_gpu_eager else None
        if sustained:
            sustained["value"] = world * 1e3 / sustained["ms_per_step"]
            sustained["unit"] = UNIT
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": f"sample-parallel x{world}",
                       "l2": "per-step working set ~3 GB (pair tensor 537 MB fp32 + workspaces) >> 126 MB L2, no flush needed",
                       "timed": "CUDA-graph replay of one full sampling step (network + DDPM update)"},
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e, "api": "ProteinReDiffModel.sample_step (graph-cached per prepared batch), pinned host buffers"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": eager,
            "sustained": sustained,
            "ragged": ragged,
            "sample_parallel": sp,
            "flops_per_step": 6730.6e9,
            "achieved_tflops_step": 6730.6e9 / (ms_per_step * 1e-3) / 1e12,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sample_parallel(model, dev, rank, world, t_steps=8, samples=64, micro=BATCH):
    """64 samples of one 512-token complex through sampling.sample_parallel_model's recipe with the real model and NCCL:
    joint prepare_batch on the full batch, rows rank::world, micro-batches of 8 through model.sample (CUDA graph per
    micro-batch), ONE final all_gather.  Strong scaling: the job is fixed, every rank samples 64 / world rows."""
    import dataclasses

    import torch.distributed as dist

    from protein_redesign_b200.sampling import shard_rows, unshard_rows

    if samples % (world * micro) != 0:
        return {"skipped": f"{samples} samples do not split into micro-batches of {micro} over {world} ranks"}
    cfg = dataclasses.replace(syn.PAPER, num_steps=t_steps, mask_prob=0.15)
    saved = (model.cfg, model.num_steps, model.mask_prob, model.setup_schedule)
    model.cfg, model.num_steps, model.mask_prob = cfg, t_steps, 0.15
    model.run_setup_schedule()
    try:
        one = syn.make_batch(cfg, [(N_ATOMS, N_TOKENS - N_ATOMS)], seed=500)
        full = {k: (v.expand(samples, *v.shape[1:]).contiguous() if isinstance(v, torch.Tensor) else v) for k, v in one.items()}
        g = torch.Generator().manual_seed(1234)
        noise = {"z_T": torch.randn(samples, N_TOKENS, 3, generator=g), "seq_T": torch.randn(samples, N_TOKENS, 21, generator=g),
                 "steps": torch.randn(t_steps - 1, samples, N_TOKENS, 3, generator=g)}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        torch.manual_seed(99)  # every rank draws the SAME joint residue mask (as every DDP rank seeds identically)
        prepared = model.prepare_batch({k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in full.items()})
        mine = shard_rows(prepared, rank, world)
        nz = {k: (v[:, rank::world] if k == "steps" else v[rank::world]) for k, v in noise.items()}
        rows = samples // world
        pos_l, log_l = [], []
        for m0 in range(0, rows, micro):
            mb = {k: (v[m0:m0 + micro].contiguous() if isinstance(v, torch.Tensor) and v.dim() >= 1 else v) for k, v in mine.items()}
            nb = {k: (v[:, m0:m0 + micro] if k == "steps" else v[m0:m0 + micro]).contiguous() for k, v in nz.items()}
            p_, l_ = model.sample(mb, noise=nb, prepared=True)
            pos_l.append(p_)
            log_l.append(l_)
        pos, logits = torch.cat(pos_l), torch.cat(log_l)
        e1.record()
        if world > 1:
            pp = [torch.empty_like(pos) for _ in range(world)]
            lp = [torch.empty_like(logits) for _ in range(world)]
            dist.all_gather(pp, pos)
            dist.all_gather(lp, logits)
            pos, logits = unshard_rows(pp, samples), unshard_rows(lp, samples)
        e2.record()
        torch.cuda.synchronize()
        total_ms, gather_us = e0.elapsed_time(e2), e1.elapsed_time(e2) * 1e3
        if world > 1:
            tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_ms = float(tt[0])
        batch_steps = samples // micro * t_steps  # batch-of-8 denoiser steps in the whole job
        return {"config": "64 samples of one 512-token complex, paper dims, micro-batches of 8, rank::world sharding",
                "samples": samples, "num_steps": t_steps, "ranks": world, "seconds": total_ms * 1e-3,
                "value": batch_steps / (total_ms * 1e-3), "unit": UNIT, "scaling": "strong",
                "includes": "H2D of the 64-row batch, joint prepare_batch, per-micro-batch graph capture, sampling, final all_gather",
                "gather_us": gather_us if world > 1 else 0.0,
                "checksum": float(pos.double().abs().sum() + logits.double().abs().sum())}
    finally:
        model.cfg, model.num_steps, model.mask_prob, model.setup_schedule = saved
        model.run_setup_schedule()


TRAIN_SIZES = [[(24, 290), (35, 212)], [(12, 333), (50, 150)], [(8, 366), (20, 250)], [(33, 270), (19, 305)],
               [(27, 180), (40, 310)], [(16, 344), (30, 225)], [(45, 200), (10, 290)], [(22, 320), (38, 240)]]


def run_train(args, dev, rank, local_rank, world, lib):
    """BASELINE config 4: training_step forward + backward (reference model.py:528-549, train.py:34-49), batch 2 per GPU,
    mask_prob 0.15, num_steps 2000, PDBbind-shaped ragged complexes (39-366 residues, 8-138 ligand atoms: rows padded to the
    batch max), data-parallel gradient all-reduce (one flat fp32 bucket, NCCL) INSIDE the timed step."""
    import dataclasses

    import torch.distributed as dist

    cfg = dataclasses.replace(syn.PAPER, mask_prob=0.15, num_steps=2000)
    model = _make_model(cfg, dev, train=True)
    # every rank gets complexes of the same sizes (different content): the scaling run then measures the all-reduce, not
    # a straggler with a longer complex; --train-sizes picks another entry of TRAIN_SIZES
    sizes = TRAIN_SIZES[args.train_sizes % len(TRAIN_SIZES)]
    host = syn.make_batch(cfg, sizes, seed=700 + rank, with_positions=True)
    host = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))
    from protein_redesign_b200.autograd import FlatGrads, TrainStepGraph, training_step_manual
    grads = FlatGrads(model).attach()
    flat = grads.flat
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    torch.manual_seed(rank)
    to_dev = lambda: {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    tsg = TrainStepGraph(model, to_dev(), grads) if args.train_mode == "graph" else None

    def step():
        batch = to_dev() if tsg is None else host   # the graph path copies the pinned batch itself, on its side stream
        if args.train_mode == "autograd":   # the drop-in path: training_step + loss.backward() through autograd
            flat.zero_()
            loss = model.training_step(batch, 0)
            loss.backward()
        elif args.train_mode == "manual":   # same kernels, gradients written straight into the flat bucket
            loss = training_step_manual(model, batch, grads)
        else:                               # device side of the step replayed as one CUDA graph
            loss = tsg.step(batch)
        if world > 1:
            dist.all_reduce(flat)
            flat.mul_(1.0 / world)
        loss_host.copy_(loss.detach().reshape(()), non_blocking=True)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    c0 = lib.prd_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    launches = int(lib.prd_launch_count() - c0)
    if tsg is not None:   # replays do not pass through the library's launch counter: count what the capture recorded
        launches = tsg.launches_per_step * args.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    if rank == 0:
        ms_step = ms / args.steps
        line = {"metric": "training steps/sec (fwd+bwd, batch 2 per GPU, paper dims)", "value": world * 1e3 / ms_step,
                "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands forward, tf32 operands backward, f32 accumulate / gradients", "data": "synthetic",
                "config": {"workload": "training_step fwd+bwd, batch 2 per GPU, mask_prob 0.15, num_steps 2000, PDBbind-shaped "
                                       f"ragged complexes {sizes} (rank 0), gradient all-reduce of {flat.numel() * 4 / 1e6:.0f} MB fp32 inside "
                                       "the timed step", "global_batch": 2 * world, "parallelism": f"data-parallel x{world}"},
                "e2e": {"value": world * 1e3 / ms_step, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "api": {"autograd": "ProteinReDiffModel.training_step + loss.backward()", "manual": "autograd.training_step_manual",
                                "graph": "autograd.TrainStepGraph.step"}[args.train_mode] + ", pinned host batch copied every step"},
                "train_mode": args.train_mode,
                "gpu_launches": launches, "gpu_launches_per_step": launches // max(args.steps, 1), "clocks": clocks,
                "loss": float(loss_host)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


CONFIGS = {
    # BASELINE.json configs 1, 2, 5 (config 3 is the default workload, config 4 is --workload train)
    "config1": ("README", [(30, 110)] * 3, "README example dims 256/32/4 blocks, 110 residues + 30 ligand atoms, batch 3"),
    "config2": ("PAPER", [(30, 270)], "paper dims, ~300-token complex (30 + 270), batch 1"),
    "config5": ("PAPER", [(1, 1023)], "paper dims, 1024 tokens = 1023 residues + 1 dummy atom (ligand-free '*' mode), batch 1"),
}


def run_config(args, dev, rank, lib):
    """The other BASELINE.json configurations as driver-visible lines: CUDA-graph replay of one sampling step."""
    if rank != 0:
        return
    cfg_name, sizes, what = CONFIGS[args.workload]
    cfg = getattr(syn, cfg_name)
    model = _make_model(cfg, dev)
    sg = StepGraph(model, cfg, syn.make_batch(cfg, sizes, seed=100), dev, 0, lib)
    with torch.inference_mode():
        sampler = ClockSampler(dev.index or 0)
        ms = sg.time_replays(args.steps, args.warmup, None, sampler)
        clocks = sampler.stop()
    ms_step = ms / args.steps
    print(json.dumps({"metric": f"denoiser steps/sec ({what})", "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f16 operands / f32 accumulate", "data": "synthetic", "config": {"workload": what},
                      "gpu_launches": sg.launches_per_step * args.steps, "gpu_launches_per_step": sg.launches_per_step,
                      "clocks": clocks}))


def profile_dominant(lib, cfg, B, N, dev, mask, pair, model):
    """Average duration of individual kernels, each timed alone (CUDA events on the launch stream) on the
    data the last step left in the workspace.  Returns the roofline object of the dominant kernel (the
    triangle-attention core) and a list for the other kernels north_star names."""
    from protein_redesign_b200 import _lib, ops

    peaks = load_peaks()
    d = ops.make_dims(cfg, B, N, mode=2)  # bit 1: the all-valid hint the step itself passes for this workload
    ws = _lib.Workspace.reserve(dev, max(_lib.workspace_bytes(op, d) for op in _lib.OPS))
    lib.prd_profile_kernel.restype = ctypes.c_int
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def time_kernel(name, aux):
        ms = ctypes.c_float(0.0)
        rc = lib.prd_profile_kernel(name.encode(), ctypes.byref(d), ctypes.c_void_p(ws.data_ptr()),
                                    ctypes.c_size_t(ws.numel()), ctypes.c_void_p(aux.data_ptr() if aux is not None else 0),
                                    5, ctypes.byref(ms), stream)
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        return ms.value

    P = 4.0 * B * N * N * cfg.pair_dim  # bytes of one fp32 pair tensor
    out = []
    ncu = {}
    if (B, N) == (BATCH, N_TOKENS):  # round 1's captures, overridden by this round's for the kernels that changed
        for name in ("r01_ncu_traffic.json", "r02_ncu_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                with open(tpath) as f:
                    ncu.update(json.load(f))

    def dram(key):  # dram bytes per launch from the committed ncu --set full capture (None at other sizes)
        e = ncu.get(key)
        return (e["read_bytes"] + e["write_bytes"]) if e else None

    # each kernel is timed on the operands its own op leaves in the (shared) workspace: run that op first, on a copy of
    # the last step's pair tensor -- the attention core's lazy-rescale path would otherwise run on another op's bytes
    blk = model.Denoiser.folding_blocks[0]
    scratch = pair.clone()
    # triangle-multiplication contraction: 2*B*N^3*c_z flop; a, b fp16 planes in, x fp16 planes out (fp32 accumulate)
    blk.pair_mul_outgoing.apply_(cfg, scratch, mask)
    ms = time_kernel("trimul_gemm", None)
    flops = 2.0 * B * N ** 3 * cfg.pair_dim
    out.append({"kernel": "gemm_f16_kernel (tri-mul contraction)", "bound": "tensor", "achieved": flops / ms / 1e9,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": flops / ms / 1e9 / peaks["tflops"],
                "traffic": dram("gemm_f16_kernel<256,4> (tri-mul contraction)"),
                "ms_per_launch": ms, "hbm_floor_ms": 1.5 * P / peaks["hbm_gbs"] / 1e6,
                "hbm_frac": 1.5 * P / ms / 1e6 / peaks["hbm_gbs"],
                "note": "HBM floor: 0.5 P (a) + 0.5 P (b) + 0.5 P (x as fp16 planes) = 1.5 P; ncu tensor-pipe active 49.5 % "
                        "(profiles/r01_ncu_traffic.json: the kernel is unchanged since round 1)"})
    # pair-bias stream: reads P, writes P/16
    ms = time_kernel("pair_bias", pair)
    nbytes = P + P * 4 / cfg.pair_dim
    out.append({"kernel": "pair_bias_kernel (LN + c_z->4 bias stream)", "bound": "hbm", "achieved": nbytes / ms / 1e6,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"],
                "traffic": dram("pair_bias_kernel"),
                "ms_per_launch": ms})
    # triangle attention core (dominant): QK^T + PV flops; co-limited by 4*B*N^3 exp2 on the MUFU pipe
    blk.pair_attn_starting.apply_(cfg, scratch, mask, all_valid=True)
    ms = time_kernel("triattn_flash", mask)
    del scratch
    flops = 2.0 * 2.0 * B * N * cfg.num_heads * N * N * cfg.head_dim
    n_exp = float(B) * N * cfg.num_heads * N * N
    mufu_floor_ms = n_exp / (148 * 16 * 1.965e9) * 1e3
    traffic = dram("triattn_flash_kernel")
    roof = {"kernel": "triattn_flash_g4_kernel", "bound": "tensor", "achieved": flops / ms / 1e9, "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": flops / ms / 1e9 / peaks["tflops"], "traffic": traffic,
            "peak_source": peaks["source"], "ms_per_launch": ms, "mufu_floor_ms": mufu_floor_ms,
            "mufu_frac": mufu_floor_ms / ms,
            "hbm_achieved_gbs": (traffic / ms / 1e6) if traffic else None,
            "note": "K=16 attention (head dim 16): 275 GFLOP on the tensor pipe against %.2e exp2 per launch; the binding "
                    "unit is the exp2 path (MUFU 16/clk/SM, a quarter of the exp2 moved to an FMA-pipe polynomial), "
                    "not the tensor pipe; mufu_frac = all-MUFU floor / measured (profiles/r01_flash_variants.md, profiles/r02_launches.md)" % n_exp,
            "others": out}
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch-on-the-same-GPU baseline")
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "config1", "config2", "config5"],
                    help="sample = the headline (BASELINE config 3 shape); train = config 4 (fwd+bwd); configN = the other configs")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 300-replay / >= 8 s sustained block")
    ap.add_argument("--no-ragged", action="store_true", help="skip the ragged (10 percent padding) block")
    ap.add_argument("--no-sample-parallel", action="store_true", help="skip the real 64-sample sharded run (config 3)")
    ap.add_argument("--sp-steps", type=int, default=8, help="diffusion steps of the sample-parallel run")
    ap.add_argument("--train-mode", default="graph", choices=["autograd", "manual", "graph"],
                    help="--workload train: drop-in autograd path / manual backward into a flat bucket / the same as one CUDA graph")
    ap.add_argument("--train-sizes", type=int, default=0, help="--workload train: which entry of TRAIN_SIZES")
    ap.add_argument("--reference-b1", action="store_true", help="--impl reference: time one complex instead of the batch of 8")
    ap.add_argument("--profile-eager", action="store_true",
                    help="run the steps eagerly (no CUDA graph, no e2e / CPU legs): for ncu launch lists")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
