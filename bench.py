#!/usr/bin/env python
"""bench.py -- ViT-B/16 PAD fine-tune throughput on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

A "step" is the reference's optimisation step (/root/reference/train_advanced.py:322-346): forward,
focal loss, backward, global-norm clip 1.0, Adam step, zero_grad, LR-schedule step, accuracy -- on one
synthetic batch of 64 images per GPU (BASELINE.json configs[1]; weak scaling for N > 1 = configs[4]).

  value      img/s over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with pinned HOST inputs: H2D of images+labels and the
             D2H of every step's loss/accuracy (pkg.HostScalars: pinned, asynchronous) inside the timed region
  roofline   the dominant kernel (tcgen05 GEMM): algorithmic FLOPs of every GEMM launch / its CUDA-event
             duration, recorded live inside real steps (vitk_prof_*), against MEASURED_PEAKS.json
  cpu_baseline  the oracle (timm-semantics restatement of the reference, fp32) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ViT-B/16 fine-tune img/s"
TRAIN_FLOP_PER_IMG = 105.150e9
FWD_FLOP_PER_IMG = 35.127e9
PER_GPU_BATCH = 64


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 10 ms while running (pynvml)."""

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle (restated reference) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps: int, warmup: int, batch: int = 8, budget_s: float = 150.0):
    import torch
    from oracle import vit_oracle as vo   # the one place bench.py executes oracle/: the CPU baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = vo.OracleViTFaceAntiSpoofing(dropout=0.1, depth=12)
    vo.seeded_init_(model, seed=42)
    opt = vo.make_optimizer(model.parameters(), "adam", lr=1e-5, weight_decay=1e-4)
    crit = vo.OracleFocalLoss(0.25, 2.0)
    images, labels = vo.synthetic_batch(batch, seed=42)
    t0 = time.perf_counter()
    vo.oracle_train_step(model, crit, opt, images, labels, 1.0)
    first = time.perf_counter() - t0
    # keep the whole run inside the budget: shrink the step count, never the work per image
    total = max(1, warmup - 1) + steps
    if first * total > budget_s:
        steps = max(2, int(budget_s / first) - max(0, warmup - 1))
    for _ in range(max(0, warmup - 1)):
        vo.oracle_train_step(model, crit, opt, images, labels, 1.0)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        vo.oracle_train_step(model, crit, opt, images, labels, 1.0)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"img_s": batch / med, "s_per_step": med, "cores": cores, "steps_timed": steps, "batch": batch}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["img_s"], "unit": "img/s", "n_gpus": args.gpus,
        "steps": r["steps_timed"], "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": reference_config(r),
        "cpu_baseline": {"value": r["img_s"], "unit": "img/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps_timed']} timed steps of batch {r['batch']} (fwd+focal+bwd+clip+Adam, fp32, "
                                   "oracle port of the reference: timm is not installable here)"},
        "e2e": {"value": r["img_s"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def reference_config(r):
    """What the reference arm actually ran: the CPU-runnable case (BASELINE configs[0]), NOT configs[1]."""
    return {"workload": f"ViT-B/16 224x224 binary PAD head full fine-tune step (fwd + focal loss + bwd + clip 1.0 + Adam wd 1e-4), "
                        f"fp32, batch {r['batch']}, CPU host cores ({r['cores']} threads), synthetic data (BASELINE configs[0]: "
                        "the reference's CPU path, oracle port -- timm is not installable here)",
            "per_gpu_batch": r["batch"], "global_batch": r["batch"], "img": 224, "depth": 12, "dropout": 0.1,
            "optimizer": "Adam(lr 1e-5, wd 1e-4) + clip_grad_norm 1.0", "parallelism": "cpu", "device": "cpu"}


def load_gemm_traffic():
    """roofline.traffic: DRAM bytes per GEMM launch, from the tracked ncu summary named in the file (never a literal)."""
    path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        d = json.load(f)
    return d.get("dram_bytes_per_launch"), d.get("source")


def train_config(n_gpus, collective="none (single GPU)"):
    return {"workload": "ViT-B/16 224x224 binary PAD head full fine-tune step (fwd + focal loss + bwd + clip 1.0 + Adam "
                        "wd 1e-4), bf16, batch 64 per GPU, synthetic data (BASELINE configs[1]; N>1 = configs[4])",
            "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus, "img": 224, "depth": 12,
            "dropout": 0.1, "optimizer": "Adam(lr 1e-5, wd 1e-4) + clip_grad_norm 1.0 + cosine LR",
            "parallelism": f"dp{n_gpus}", "collective": collective,
            "l2": "per-step working set (~5 GB activations + 1 GB params/grads/moments) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import vit_spoof_detection_pda_b200 as pkg
    from vit_spoof_detection_pda_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # SMs NCCL may take for the gradient all-reduce; DataParallel shrinks the persistent GEMM grids by the same
        # number while a collective is in flight (see dp.py)
        # measured on this pool (gpurun_out/w_scale.log, z_scale.log): 2 GPUs 8 CTAs 93.6 %; 8 GPUs 8 / 16 / 32 CTAs
        # 85.8 / 92.7 / 94.1 % of 8 x the single-GPU rate -- the ring needs more CTAs as it grows
        os.environ.setdefault("NCCL_MAX_CTAS", "8" if world <= 2 else ("16" if world <= 4 else "32"))
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = L.load()   # fails loudly when libvitk.so is missing

    torch.manual_seed(42)
    model = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev)
    model.train()
    # VITK_GRAD_COMM=bf16 (opt-in, not the reported configuration): 16-bit gradient all-reduce, torch DDP's bf16_compress_hook
    comm_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(os.environ.get("VITK_GRAD_COMM", ""), None)
    # VITK_DP_MODE: auto (default) = the optimizer step fused with its collectives over NVSwitch multicast when the system has
    # NVLS, else the bucketed NCCL all-reduce; nvls / nccl force one
    net = pkg.DataParallel(model, bucket_mb=float(os.environ.get("VITK_BUCKET_MB", "50")), grad_comm_dtype=comm_dtype,
                           mode=os.environ.get("VITK_DP_MODE", "auto")) if world > 1 else model
    dp_mode = net.mode if world > 1 else "single"
    crit = pkg.FocalLoss(alpha=0.25, gamma=2.0)
    # single GPU: the step is captured into a CUDA graph (pkg.GraphedTrainStep; VITK_GRAPH=0 keeps the eager launches); the
    # data-parallel modes use the eager step (their collectives / hooks run between the backward stages)
    use_graph = world == 1 and os.environ.get("VITK_GRAPH", "1") != "0"
    opt = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False, capturable=use_graph)
    total_sched_steps = 2 * (args.warmup + args.steps) + 64
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=total_sched_steps, eta_min=1e-6)

    B = PER_GPU_BATCH
    g = torch.Generator().manual_seed(1234 + rank)
    n_pool = 4
    host_imgs = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(n_pool)]
    host_lbls = [torch.randint(0, 2, (B,), generator=g).pin_memory() for _ in range(n_pool)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_lbls = [t.to(dev) for t in host_lbls]

    def eager_step(images, labels):
        out = net(images)
        loss, met = crit(out, labels, with_metrics=True)
        loss.backward()
        pkg.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        sched.step()
        return loss, met

    launches_per_step = None
    if use_graph:
        eager_step(dev_imgs[0], dev_lbls[0])                   # also counts the kernels one step launches
        torch.cuda.synchronize()
        l0 = lib.vitk_launch_count()
        eager_step(dev_imgs[1], dev_lbls[1])
        launches_per_step = lib.vitk_launch_count() - l0
        gstep = pkg.GraphedTrainStep(model, crit, opt, dev_imgs[0], dev_lbls[0], max_grad_norm=1.0)

        def step(images, labels):
            loss, met = gstep(images, labels)
            sched.step()
            return loss, met
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.vitk_launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.vitk_launch_count() - l0
        if launches_per_step is not None:
            launches = launches_per_step * steps             # graph replays launch the captured kernels
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # ---- device-resident throughput ("value")
    for i in range(args.warmup):
        step(dev_imgs[i % n_pool], dev_lbls[i % n_pool])
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total, launches = timed(lambda i: step(dev_imgs[i % n_pool], dev_lbls[i % n_pool]), args.steps)
    sampler.stop()
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- sustained rate: the same step over a >= 3 s timed region (the 20-step `value` region is ~0.2 s: on power-capped
    # boxes the SM clock sags after that), with its own clock summary.  Reported in `extra`, not as `value`.
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(3300.0 / (ms_total / args.steps)))
        sus_sampler = ClockSampler(local_rank)
        sus_sampler.start()
        ms_sus, _ = timed(lambda i: step(dev_imgs[i % n_pool], dev_lbls[i % n_pool]), n_sus)
        sus_sampler.stop()
        sustained = {"steps": n_sus, "ms_per_step": ms_sus / n_sus, "img_s": world * B * n_sus / (ms_sus / 1e3),
                     "clocks": sus_sampler.summary()}

    eager_ms = None
    if use_graph:       # the same step with eager launches, for the record (extra.eager_ms_per_step)
        ms_eager, _ = timed(lambda i: eager_step(dev_imgs[i % n_pool], dev_lbls[i % n_pool]), max(10, args.steps // 2))
        eager_ms = ms_eager / max(10, args.steps // 2)

    # ---- end to end through the public API with host buffers ("e2e"): every step copies its own batch from pinned
    # host memory (DevicePrefetcher: the copy of batch i+1 overlaps step i) and reads loss + accuracy back
    def host_batches(n):
        for i in range(n):
            yield host_imgs[i % n_pool], host_lbls[i % n_pool]

    # loss and accuracy of EVERY step reach the host (the reference's two reads, train_advanced.py:345-346), through
    # pkg.HostScalars: async copies into pinned memory, values handed over one step late, so the compute stream never
    # drains (a plain .item() pair costs +0.76 ms per step: tools/e2e_probe.py)
    reader = pkg.HostScalars(dev)

    def e2e_loop(n):
        out = None
        for images, labels in pkg.DevicePrefetcher(host_batches(n), dev):
            loss, met = step(images, labels)
            prev = reader.push(loss, met["ncorrect"])
            if prev is not None:
                out = prev
        last = reader.flush()          # the final step's values: read inside the timed region too
        return last if last is not None else out

    e2e_loop(2)
    barrier()
    t_e0, t_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e0.record()
    e2e_loop(args.steps)
    t_e1.record()
    barrier()
    ms_e2e = t_e0.elapsed_time(t_e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * 3 * 224 * 224 * 4 + B * 8
    d2h = 8 + 8   # loss and ncorrect as float64 scalars

    # ---- live per-launch GEMM timing inside real steps -> roofline of the dominant kernel.  Every rank runs the two
    # steps (they contain the gradient all-reduce); only rank 0 records and reads the events.
    # (per-launch events only mean something when kernels do not overlap: the weight-gradient side stream is switched
    #  off for these two profiled steps -- vitk_model.flags = VITK_FLAG_WGRAD_INLINE -- and back on afterwards)
    # Two un-profiled eager steps are queued first and NOT waited for: the host enqueues a step in ~4 ms, the GPU runs it in
    # ~9 ms, so the profiled launches are enqueued several ms before the GPU reaches them and no event pair can contain a
    # host-side stall (the first eager launches after the graph replays, event creation).
    model._flags = L.FLAG_WGRAD_INLINE
    for i in range(2):
        eager_step(dev_imgs[i % n_pool], dev_lbls[i % n_pool])
    if rank == 0:
        lib.vitk_prof_enable(1)
    for i in range(2):
        eager_step(dev_imgs[i % n_pool], dev_lbls[i % n_pool])      # per-launch events need the eager launches
    barrier()
    model._flags = 0

    if world > 1 and dp_mode == "nvls" and os.environ.get("VITK_NVLS_PROF") == "1":
        rep = model._nvls.report()
        sys.stderr.write(f"[rank {rank}] NVLS step phases (ms): {rep}\n")
    line = None
    if rank == 0:
        peaks = load_peaks()
        import ctypes as C
        maxn = 4096
        ms_arr = (C.c_float * maxn)()
        info = (C.c_int * (5 * maxn))()
        n = lib.vitk_prof_read(ms_arr, info, maxn)
        lib.vitk_prof_enable(0)
        fl, tt, fam = 0.0, 0.0, {}
        traffic_bytes, traffic_src = load_gemm_traffic()
        algo_bytes = 0.0
        for k in range(n):
            I, J, R, mode, eng = info[5 * k:5 * k + 5]
            if eng != L.ENGINE_TCGEN05:
                continue
            f = 2.0 * I * J * R
            # operands bf16 once; output bf16 (fp32 for the residual / accumulate epilogues, two outputs for GELU)
            algo_bytes += 2.0 * R * (I + J) + I * J * {1: 4.0, 2: 8.0, 5: 4.0, 4: 4.0, 6: 4.0}.get(mode, 2.0)
            fl += f
            tt += ms_arr[k] * 1e-3
            key = f"{I}x{J}x{R}/epi{mode}"
            a = fam.setdefault(key, [0, 0.0, 0.0, []])
            a[0] += 1; a[1] += f; a[2] += ms_arr[k] * 1e-3; a[3].append(ms_arr[k] * 1e3)
        achieved = fl / tt / 1e12 if tt > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "vitk::gemm_tc_kernel (tcgen05.mma kind::f16, all GEMM launches of a step)",
                "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"],
                # dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, averaged over the launches of one step, from
                # the tracked ncu summary profiles/gemm_traffic.json names (null when that file is absent)
                "traffic": traffic_bytes, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": algo_bytes / max(1, n),
                "peak_source": peaks["source"] +
                " (sustained cuBLAS bf16: kernel timed inside a long step)", "gemm_launches_per_step": n // 2,
                "gemm_share_of_step": (tt / 2) / (ms_total / args.steps / 1e3)}
        fam_sorted = sorted(fam.items(), key=lambda kv: -kv[1][2])
        sys.stderr.write("GEMM families (2 profiled steps): shape IxJxR/epilogue  launches  TFLOP/s  ms total  us/launch min / median / max\n")
        for key, (cnt, f, t, us) in fam_sorted:
            us.sort()
            sys.stderr.write(f"  {key:32s} {cnt:4d} {f / t / 1e12:8.1f} {t * 1e3:8.3f}   {us[0]:7.1f} / {us[len(us) // 2]:7.1f} / {us[-1]:7.1f}\n")

        # ---- batch-1 latency (CUDA-graph replay) and batch-256 throughput of the eval path (config 3)
        extra = {}
        try:
            if args.no_extras:
                raise RuntimeError("skipped (--no-extras)")
            model.eval()
            x1 = dev_imgs[0][:1].clone()
            with torch.no_grad():
                for _ in range(3):
                    model(x1)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    y1 = model(x1)
                lat = []
                for _ in range(1000):      # >= 1,000 iterations, p50 / p90 / p99
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); graph.replay(); e1.record(); e1.synchronize()
                    lat.append(e0.elapsed_time(e1))
                lat.sort()
                extra["bs1_latency_iters"] = len(lat)
                extra["bs1_latency_ms_p50"] = lat[len(lat) // 2]
                extra["bs1_latency_ms_p90"] = lat[int(len(lat) * 0.90)]
                extra["bs1_latency_ms_p99"] = lat[int(len(lat) * 0.99)]
                x256 = torch.randn(256, 3, 224, 224, device=dev)
                for _ in range(2):
                    model(x256)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    model(x256)
                e1.record(); torch.cuda.synchronize()
                extra["bs256_infer_img_s"] = 256 * 5 / (e0.elapsed_time(e1) / 1e3)
            model.train()
            # ---- BASELINE configs[3]: frozen backbone (forward-only encoder + head backward + Adam on the head), batch 256
            fm = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev)
            fm.load_state_dict(model.state_dict())
            for prm in fm.vit.parameters():
                prm.requires_grad_(False)
            fm.train()
            fopt = pkg.FusedAdam(fm.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False)
            y256 = torch.randint(0, 2, (256,), device=dev)

            def fstep():
                floss = crit(fm(x256), y256)
                floss.backward()
                pkg.clip_grad_norm_(fm.parameters(), 1.0)
                fopt.step()
                fopt.zero_grad(set_to_none=True)

            for _ in range(2):
                fstep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fstep()
            e1.record(); torch.cuda.synchronize()
            extra["frozen_backbone_bs256_img_s"] = 256 * 5 / (e0.elapsed_time(e1) / 1e3)
            del fm, fopt
        except Exception as e:  # noqa: BLE001 -- extras never invalidate the headline line
            extra["eval_extra_error"] = repr(e)

        # ---- HBM-bound kernels against the measured copy bandwidth, CUDA-event timed here: 20 back-to-back launches on
        # rotating bs-64 buffers (4 copies x 39..155 MB > the 126 MB L2); Adam over the model's own flat buffers
        try:
            if args.no_extras:
                raise RuntimeError("skipped (--no-extras)")
            extra["hbm_kernels"] = hbm_kernel_rooflines(torch, L, model, opt, dev, peaks["hbm_gbs"])
        except Exception as e:  # noqa: BLE001
            extra["hbm_kernels_error"] = repr(e)
        if sustained is not None:
            extra["sustained"] = sustained
        # ---- the practical bar on the same GPU in the same run: the reference's step written with stock PyTorch library
        # kernels only (tools/torch_eager_baseline.py; nothing of this package or oracle/ on that path)
        if world == 1 and not args.no_eager_baseline:
            try:
                del x256
                torch.cuda.empty_cache()
                from tools import torch_eager_baseline as teb
                eb = teb.run(batch=B, steps=10, warmup=3, with_inference=False, device=str(dev))
                extra["torch_eager_same_gpu"] = eb
                extra["speedup_vs_torch_eager"] = value / eb["img_s"]
            except Exception as e:  # noqa: BLE001
                extra["torch_eager_error"] = repr(e)

        extra["step_mode"] = "cuda graph replay (pkg.GraphedTrainStep)" if use_graph else "eager launches"
        if eager_ms is not None:
            extra["eager_ms_per_step"] = eager_ms
        per_gpu = value / world
        extra.update({
            "tflops_per_gpu": per_gpu * TRAIN_FLOP_PER_IMG / 1e12,
            "mfu_vs_measured_sustained": per_gpu * TRAIN_FLOP_PER_IMG / 1e12 / peaks["bf16_tflops_sustained"],
            "mfu_vs_measured_burst": per_gpu * TRAIN_FLOP_PER_IMG / 1e12 / peaks["bf16_tflops"],
            "mfu_vs_spec_2250": per_gpu * TRAIN_FLOP_PER_IMG / 1e12 / 2250.0,
        })

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(steps=3, warmup=1, budget_s=60.0)
            cpu = {"value": r["img_s"], "unit": "img/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{r['steps_timed']} timed steps of batch {r['batch']} (BASELINE configs[0]: fwd+focal+bwd+clip+Adam, fp32)"}

        line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": train_config(world, {
                    "single": "none (single GPU)",
                    "nvls": "gradient reduce-scatter (multimem.ld_reduce) + Adam + parameter all-gather (multimem.st) fused in two "
                            "kernels over NVSwitch multicast, after backward",
                    "nccl": "bucketed fp32 ncclAllReduce (AVG) overlapped with backward, NCCL_MAX_CTAS=" + os.environ.get("NCCL_MAX_CTAS", "?"),
                }[dp_mode]), "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "extra": extra}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


def hbm_kernel_rooflines(torch, L, model, opt, dev, hbm_gbs):
    """LayerNorm forward / backward at the bench shape (12,608 x 768 rows) and the fused Adam pass: algorithmic bytes
     / CUDA-event time, as a fraction of the measured copy bandwidth."""
    M, D, reps, ncopy = PER_GPU_BATCH * 197, 768, 20, 4
    st = L.stream_ptr()
    x = [torch.randn(M, D, device=dev) for _ in range(ncopy)]
    y16 = [torch.empty(M, D, dtype=torch.bfloat16, device=dev) for _ in range(ncopy)]
    dx = [torch.randn(M, D, device=dev) for _ in range(ncopy)]
    dx16 = [torch.empty(M, D, dtype=torch.bfloat16, device=dev) for _ in range(ncopy)]
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    dg, db, cs = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)

    def ev_time(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    def ln_fwd(i):
        k = i % ncopy
        L.call("vitk_layernorm_fwd", L.ptr(x[k]), D, L.ptr(gamma), L.ptr(beta), L.ptr(y16[k]), L.BF16, L.ptr(mean), L.ptr(rstd),
               M, 1e-6, st)

    def ln_bwd(i):
        k = i % ncopy
        L.call("vitk_layernorm_bwd", L.ptr(y16[k]), L.BF16, L.ptr(x[k]), D, L.ptr(gamma), L.ptr(mean), L.ptr(rstd), L.ptr(dx[k]),
               L.ptr(dx[k]), L.ptr(dx16[k]), L.ptr(dg), L.ptr(db), L.ptr(cs), M, st)

    out = {}
    t = ev_time(ln_fwd)
    b = M * D * (4 + 2) + M * 8
    out["ln_fwd_kernel"] = {"bound": "hbm", "bytes": b, "us": t * 1e6, "achieved": b / t / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                            "frac": b / t / 1e9 / hbm_gbs}
    ln_fwd(0)
    t = ev_time(ln_bwd)
    b = M * D * (2 + 4 + 4 + 4 + 2)
    out["ln_bwd_kernel"] = {"bound": "hbm", "bytes": b, "us": t * 1e6, "achieved": b / t / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                            "frac": b / t / 1e9 / hbm_gbs}
    del x, y16, dx, dx16
    # Adam: one pass over p, g, m, v (+ bf16 shadow): 30 B / parameter.  lr = 0 so the timed passes do not move the weights.
    flat, g = model.flat_params(), model.flat_grads()
    n = flat.numel()
    p16 = model.flat_params16()
    bm, bv = torch.zeros_like(flat), torch.zeros_like(flat)      # scratch moments: the optimizer's own state is not touched

    def adam(i):
        L.call("vitk_adam_step", L.ptr(flat), L.ptr(g), L.ptr(bm), L.ptr(bv), L.ptr(p16), n, 0.0, 0.9, 0.999, 1e-8, 0.0, 0,
               1000, 1.0, None, 0.0, st)

    t = ev_time(adam)
    del bm, bv
    b = n * 30
    out["adam_kernel"] = {"bound": "hbm", "bytes": b, "us": t * 1e6, "achieved": b / t / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                          "frac": b / t / 1e9 / hbm_gbs}
    return out


_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything libraries print to fd 1 (e.g. NCCL's version banner) to stderr, keeping the real stdout for
    the single JSON line the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the inference / frozen-backbone / HBM-kernel extras (scaling sweeps)")
    ap.add_argument("--no-eager-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    run_ours(args)


if __name__ == "__main__":
    main()
