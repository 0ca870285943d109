"""bench.py -- objects/sec of the SwinVox multi-view reconstruction forward (3 views 224x224 -> 32^3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--views V]

One "step" = one pass of the hot path over one batch of synthetic input: encoder -> decoder -> merger -> refiner ->
threshold/IoU counters (core/test.py:120-164) on B objects x V views per GPU (BASELINE.json configs[1]: batch 64 x
3 views, merger + refiner, fp32/TF32).  N>1 is launched by torchrun, one rank per GPU; objects are sharded (weak
scaling: B per GPU fixed) and the only exchange is the NCCL all_gather of logits + counters inside the timed region.
`--impl reference` times the reference's CPU forward (the oracle port, bit-exact to /root/reference/models/*.py)
on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "objects/sec, 3 views 224^2 -> 32^3"
# algorithmic FLOPs (2*MAC) of the reference forward per object: 19.382*V + 2.485 GFLOP with CVA on
GF_PER_VIEW, GF_PER_OBJECT = 19.382, 2.485
GF_ATTENTION_PER_VIEW = 0.280 + 0.0561   # window-attention bmm + CVA: run outside the contraction kernel
GF_MERGER_PER_VIEW = 1.1625              # merger convolutions: conv3_slab_kernel, not the dominant kernel
GF_ENCODER_PER_VIEW = 17.833             # encoder total per view (the part that runs in bf16 with --dtype bf16)
TF32_CUBLAS_MEASURED = 712.5             # torch.matmul fp32/TF32 8192^3 on this pool (profiles/r1_gemm_bench_v6.log)


def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU during the timed region (NVML)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def workload_tag(B, V, world, bf16):
    """which BASELINE.json configuration a (batch per GPU, views, ranks, dtype) combination is"""
    if bf16:
        tag = "BASELINE configs[2]" if (B, V) == (64, 5) else "bf16 variant"
        return f"({tag}; encoder in bf16, decoder / merger / refiner in TF32)"
    if (B, V) == (64, 3):
        return "(BASELINE configs[1])"
    if V == 20 and B * world == 256:
        return "(BASELINE configs[3]: batch 256 x 20 views over the ranks)"
    if B * world == 128:
        return "(BASELINE configs[4]: a point of the 1-24 view sweep at batch 128)"
    return "(a shape outside BASELINE.json's list)"


def synthetic_batch(B, V, seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, V, 3, 224, 224, generator=g) * 2 - 1
    gt = (torch.rand(B, 32, 32, 32, generator=g) < 0.1).float()
    return images, gt


def cpu_reference_rate(V, seconds_budget, sample_objects=2):
    """reference CPU forward (oracle port) on all host cores; returns (objects/sec, cores, description)"""
    from oracle import modules as M
    torch.set_num_threads(os.cpu_count())
    cfg = M.default_cfg()
    torch.manual_seed(0)
    enc, dec, mer, ref = M.RefEncoder(cfg).eval(), M.RefDecoder(cfg).eval(), M.RefMerger(cfg).eval(), M.RefRefiner(cfg).eval()
    images, gt = synthetic_batch(sample_objects, V, 1)
    with torch.no_grad():
        def step():
            vol = M.forward_pipeline(enc, dec, mer, ref, images, cfg)
            return M.voxel_metrics(vol, gt)
        step()  # warm-up
        t0, n = time.perf_counter(), 0
        while True:
            step()
            n += 1
            if time.perf_counter() - t0 > seconds_budget or n >= 50:
                break
        dt = (time.perf_counter() - t0) / n
    return sample_objects / dt, os.cpu_count(), f"{n} timed passes of {sample_objects} objects x {V} views (after 1 warm-up)"


def gpu_eager_rate(B, V, dev, seconds_budget=8.0):
    """The on-box bar: the reference's modules as eager PyTorch on the SAME B200 -- cuDNN / cuBLAS,
    cudnn.benchmark on (core/test.py:35), inputs resident, no per-sample host syncs, batched B x V (kinder than the
    reference's batch-1 loop).  Timed twice: PyTorch's default precision (TF32 convolutions, fp32 matmul = what the
    reference runs) and with TF32 matmul allowed as well.  Baseline leg only: nothing of it is on the product path."""
    from oracle import modules as M
    cfg = M.default_cfg()
    torch.manual_seed(0)
    mods = [M.RefEncoder(cfg), M.RefDecoder(cfg), M.RefMerger(cfg), M.RefRefiner(cfg)]
    enc, dec, mer, ref = [m.eval().to(dev) for m in mods]
    images, gt = synthetic_batch(B, V, 1)
    images, gt = images.to(dev), gt.to(dev)
    saved = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32)
    out = {}
    try:
        torch.backends.cudnn.benchmark = True
        for key, mm_tf32 in (("value", False), ("value_tf32_matmul", True)):
            torch.backends.cuda.matmul.allow_tf32 = mm_tf32
            with torch.no_grad():
                def step():
                    vol = M.forward_pipeline(enc, dec, mer, ref, images, cfg)
                    return M.voxel_metrics(vol, gt)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n, t0 = 0, time.perf_counter()
                e0.record()
                while n < 3 or (time.perf_counter() - t0 < seconds_budget / 2 and n < 30):
                    step()
                    n += 1
                e1.record()
                torch.cuda.synchronize()
            out[key] = B / (e0.elapsed_time(e1) / n * 1e-3)
            out[key + "_steps"] = n
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32 = saved
    del enc, dec, mer, ref, mods
    torch.cuda.empty_cache()
    out.update(unit="objects/s", kind="port (oracle modules, bit-exact to the reference's, as eager PyTorch on cuda)",
               sample=f"batch {B} x {V} views resident on the device, cudnn.benchmark on; value = PyTorch default "
                      "precision (TF32 convolutions, fp32 matmul), value_tf32_matmul = TF32 matmul allowed too")
    return out


def run_reference_gpu(args):
    """`--impl reference-gpu`: the eager-PyTorch-on-B200 bar alone, as its own JSON line (N=1 only)"""
    rank, local_rank, world = rank_env()
    if rank != 0:
        return
    dev = torch.device("cuda", local_rank)
    r = gpu_eager_rate(args.batch, args.views, dev, seconds_budget=max(8.0, args.steps))
    line = {"impl": "reference-gpu", "metric": METRIC, "value": r["value"], "unit": "objects/s", "n_gpus": 1,
            "steps": r["value_steps"], "warmup": 3, "ms_per_step": args.batch / r["value"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 convolutions / fp32 matmul (PyTorch default)",
            "data": "synthetic",
            "config": {"workload": f"batch {args.batch} x {args.views} views, merger + refiner, CVA on (BASELINE configs[1])"},
            "gpu_eager_baseline": r}
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank, _, world = rank_env()
    if rank != 0:
        return
    from oracle import modules as M
    torch.set_num_threads(os.cpu_count())
    cfg = M.default_cfg()
    torch.manual_seed(0)
    enc, dec, mer, ref = M.RefEncoder(cfg).eval(), M.RefDecoder(cfg).eval(), M.RefMerger(cfg).eval(), M.RefRefiner(cfg).eval()
    sample = args.ref_sample
    images, gt = synthetic_batch(sample, args.views, 1)
    with torch.no_grad():
        def step():
            vol = M.forward_pipeline(enc, dec, mer, ref, images, cfg)
            return M.voxel_metrics(vol, gt)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "objects/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch {args.batch} x {args.views} views, merger + refiner, CVA on (BASELINE configs[1])",
                   "step": f"bounded sample: {sample} objects x {args.views} views per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": "objects/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} steps of {sample} objects x {args.views} views, {os.cpu_count()} threads"},
        "e2e": {"value": value, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from swinvox_b200 import config as svx_config
    from swinvox_b200.metrics import VoxelMetrics
    from swinvox_b200.pipeline import DataParallelReconstructor, Reconstructor
    rank, local_rank, world = rank_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (swinvox_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, V = args.batch, args.views
    cfg = svx_config.make_cfg()
    torch.manual_seed(0)
    bf16 = args.dtype == "bf16"
    # random-init weights of the reference architecture; --dtype bf16: the encoder stores / multiplies in bf16
    rec = Reconstructor(cfg, device=dev, zero_copy=True, dtype=args.dtype)
    rec.set_graph(not args.no_graph)
    dp = DataParallelReconstructor(rec)
    images_h, gt_h = synthetic_batch(B, V, 100 + rank)
    images_h, gt_h = images_h.pin_memory(), gt_h.pin_memory()
    inbuf = rec.input_buffer(B, V)
    inbuf.copy_(images_h)
    gt_d = gt_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        # N > 1: the gather of step i runs on NCCL's stream under the forward of step i+1 (drained by dp.flush() below)
        return dp.evaluate_local(inbuf, gt_d, wait=False)

    # ---- device-resident throughput ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    dp.flush()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    dp.flush()   # every step's gathered logits + counters have arrived before the clock stops
    e1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- end to end through the public evaluation driver (swinvox_b200.evaluate.BatchedEvaluator): every step copies
    # its images from pinned host memory (copy stream, overlapped with the previous step's kernels) and reads its
    # logits + counters back to pinned host memory (output stream); per-taxonomy IoU accumulates on the device --------
    from swinvox_b200.evaluate import BatchedEvaluator
    got = {"batches": 0, "voxels": 0.0}

    def consume(logits_host, counts_host):   # the user's per-batch consumer: touches the host copies
        got["batches"] += 1
        got["voxels"] += float(counts_host[:, -1, 0].sum())

    tax = ["synthetic"] * B
    ev = BatchedEvaluator(rec, B, V, on_batch=consume)
    for _ in range(2):
        ev.submit(tax, images_h, gt_h)
    ev.finish(print_tables=False)
    barrier()
    ev = BatchedEvaluator(rec, B, V, on_batch=consume)
    got["batches"] = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev.submit(tax, images_h, gt_h)
    max_iou, report = ev.finish(print_tables=False)
    barrier()
    # finish() adds the per-taxonomy sums over the ranks (one all_reduce): the report counts the whole job's samples
    assert got["batches"] == args.steps and report["n_samples"] == B * args.steps * world, (got, report["n_samples"])
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (e2e_s.item() / args.steps)
    h2d_bytes = images_h.numel() * 4 + gt_h.numel() * 4
    d2h_bytes = B * 32768 * 4 + B * len(cfg.TEST.VOXEL_THRESH) * 5 * 4

    # ---- roofline of the dominant kernel (the tcgen05 contraction kernel), measured live with CUDA events -------
    gemm_ms, slab_ms, mlp_ms, mlp_gf, total_ms, gemm_bytes, breakdown, n_gemm, other_gemm_ms = 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, [], 0, 0.0
    for mod in rec.modules():
        for entry in mod._plans.values():
            plan = entry[0]
            plan_bf16 = plan.dtype == torch.bfloat16
            # per-op CUDA-event times: mean of 3 launches, measured twice, the smaller mean kept (one run in v44 had a single
            # op of the list at 3x its usual time; the step time itself is not derived from these)
            times = [min(a, b) for a, b in zip(plan.time_ops(iters=3), plan.time_ops(iters=3))]
            for nm, t, fl, nb in zip(plan.op_names, times, plan.flops, plan.bytes):
                total_ms += t
                if fl > 0 and nm.startswith("merger.layer"):
                    slab_ms += t
                elif nm.endswith(".mlp"):      # mlp_fused_kernel (Swin stages 0 / 1): its own kernel, not the dominant one
                    mlp_ms += t
                    mlp_gf += fl / 1e9
                elif fl > 0 and not nm.endswith(".attn"):
                    if plan_bf16 == bf16:       # the dominant kernel: gemm_bf16_kernel (encoder) with --dtype bf16, else gemm_tf32_kernel
                        gemm_ms += t
                        gemm_bytes += nb
                        n_gemm += 1
                    else:
                        other_gemm_ms += t
                breakdown.append((nm, t, fl, nb, plan_bf16))
    # dominant kernel = gemm_tf32_kernel (every Linear / Conv2d / Conv3d k4 / ConvTranspose3d).  Algorithmic FLOPs per
    # step = SURVEY 8(d) figure of the reference forward minus what other kernels execute (attention, merger convs).
    if bf16:   # gemm_bf16_kernel executes the encoder's contractions (everything but window attention / CVA's view attention)
        algo_gf = B * V * (GF_ENCODER_PER_VIEW - GF_ATTENTION_PER_VIEW)
    else:
        algo_gf = B * ((GF_PER_VIEW - GF_ATTENTION_PER_VIEW - GF_MERGER_PER_VIEW) * V + GF_PER_OBJECT) - mlp_gf
    achieved = algo_gf / gemm_ms if gemm_ms > 0 else 0.0   # GFLOP / ms = TFLOP/s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    tf32_peak = bf16_peak / 2.0   # kind::tf32 runs at half the bf16 rate; MEASURED_PEAKS.json has no tf32 entry
    peak = bf16_peak if bf16 else tf32_peak
    hbm = peaks.get("hbm_gbs", 6550.7)
    # per-op roofline floor: every op is bounded by max(algorithmic FLOP / tensor peak, algorithmic bytes / HBM peak)
    floor_ms = sum(max(fl / ((bf16_peak if pb else tf32_peak) * 1e9), nb / (hbm * 1e6)) for nm, t, fl, nb, pb in breakdown)
    traffic = None   # DRAM bytes per launch of the dominant kernel from the committed ncu launch list of this command
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[
            "gemm_bf16_kernel" if bf16 else "gemm_tf32_kernel"]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "gemm_bf16_kernel" if bf16 else "gemm_tf32_kernel", "launches_per_step": n_gemm,
                "algorithmic_bytes_per_launch": gemm_bytes / max(n_gemm, 1),
                "step_floor_ms": floor_ms, "step_frac_of_floor": floor_ms / total_ms if total_ms > 0 else None,
                "kernel_ms_per_step": gemm_ms, "all_kernels_ms_per_step": total_ms,
                "algorithmic_gflop_per_step": algo_gf, "conv3_slab_ms_per_step": slab_ms,
                "mlp_fused_ms_per_step": mlp_ms, "mlp_fused_tflops": (mlp_gf / mlp_ms if mlp_ms > 0 else None),
                "tf32_cublas_tflops_measured": TF32_CUBLAS_MEASURED,
                "other_gemm_ms_per_step": other_gemm_ms,
                "peak_source": (("MEASURED_PEAKS.json bf16_tflops_sustained" if bf16 else
                                 "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32)") + ", of measured" if peaks
                                else "fallback 1.4 PFLOP/s sustained bf16" + ("" if bf16 else " / 2 (tf32)") + ", of fallback")}

    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "op_breakdown.json"), "w") as fh:
            json.dump(sorted([list(r[:4]) for r in breakdown], key=lambda r: -r[1]), fh)
        cpu_value, cores, sample = cpu_reference_rate(V, args.cpu_seconds) if world == 1 else (None, None, None)
        eager = None
        if world == 1 and not args.no_eager:
            try:
                eager = gpu_eager_rate(B, V, dev)
            except Exception as e:  # noqa: BLE001  (a baseline leg must never take the product's line down)
                eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": "objects/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if bf16 else "tf32", "data": "synthetic",
            "config": {"workload": f"batch {B} x {V} views per GPU, merger + refiner, CVA on, 224x224 -> 32^3 "
                                   + workload_tag(B, V, world, bf16),
                       "l2": "per-step working set (activations) is several GB, far larger than the 126 MB L2",
                       "cuda_graph": not args.no_graph, "weights": "random init (reference architecture)"},
            "e2e": {"value": e2e_value, "unit": "objects/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "api": "swinvox_b200.evaluate.BatchedEvaluator.submit/finish"},
            "gpu_launches": rec.num_launches() * args.steps,
            "clocks": sampler.summary(),
            "roofline": roofline,
        }
        if cpu_value is not None:
            line["cpu_baseline"] = {"value": cpu_value, "unit": "objects/s", "cores": cores, "kind": "port", "sample": sample}
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline leg")
    ap.add_argument("--batch", type=int, default=64, help="objects per GPU")
    ap.add_argument("--views", type=int, default=3)
    ap.add_argument("--dtype", default="tf32", choices=["tf32", "bf16"],
                    help="bf16: the encoder stores / multiplies in bf16 (BASELINE configs[2] with --views 5)")
    ap.add_argument("--ref-sample", type=int, default=4, help="objects per reference-arm step (bounded CPU sample)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline budget inside the default run")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
