#!/usr/bin/env python
"""Benchmark of the SViT hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, bf16, batch 64 clips / GPU)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU oracle port on the host cores

One "step" = one forward of SViT (configs/ssv2.yaml geometry, random init) over one batch of synthetic
16x224^2 clips.  `value` = clips/s with the batch resident in HBM; `e2e` = clips/s through the public module
call with pinned-host inputs (H2D of the clips and D2H of the class probabilities inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            d["_source"] = "measured"
            return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ work model
def attn_flops(B, spec, thw, O_tot):
    """4*B*h*Nq*Nk*96 + 2*B*h*Lq*96*(kh+kw+kt)  (SURVEY.md 8d)."""
    T, H, W = thw
    sq, skv = spec["stride_q"][1], spec["stride_kv"][1]
    qh, qw = (H - 1) // sq + 1, (W - 1) // sq + 1
    kh, kw = (H - 1) // skv + 1, (W - 1) // skv + 1
    Lq, Lk = T * qh * qw, T * kh * kw
    Nq, Nk = 1 + Lq + O_tot, 1 + Lk + O_tot
    h = spec["num_heads"]
    return 4.0 * B * h * Nq * Nk * 96 + 2.0 * B * h * Lq * 96 * (kh + kw + T), [T, qh, qw], Nq, Nk


def work_per_clip(cfg):
    """Algorithmic FLOPs / bytes per clip for each kernel family (forward)."""
    from svit_b200.config import block_specs
    specs, patch_dims, _ = block_specs(cfg)
    O_tot = cfg.DATA.NUM_FRAMES * cfg.SVIT.O
    thw = list(patch_dims)
    L = thw[0] * thw[1] * thw[2]
    fl = {"attn": 0.0, "gemm": 2.0 * L * 96 * 441, "pool_bytes_bf16": 0.0}
    for sp in specs:
        N = 1 + thw[0] * thw[1] * thw[2] + O_tot
        C, D, h = sp["dim"], sp["dim_out"], sp["num_heads"]
        f, q_thw, Nq, Nk = attn_flops(1, sp, thw, O_tot)
        fl["attn"] += f
        fl["gemm"] += 2.0 * N * C * 3 * D + 2.0 * Nq * D * D + 2 * 2.0 * Nq * D * 4 * D
        if C != D:
            fl["gemm"] += 2.0 * N * C * D
        fl["pool_bytes_bf16"] += 2.0 * 96 * h * ((N + Nq) + 2 * (N + Nk))
        thw = q_thw
    return fl


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_forward_timing(max_seconds=25.0, max_runs=5):
    """Times the CPU oracle port of the reference forward (fp32, batch 1, all host threads)."""
    from oracle import svit_oracle as O
    from svit_b200.config import block_specs, ssv2_cfg, state_shapes
    from tests.golden.recipe import synth_input, synth_state

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ssv2_cfg()
    params = synth_state(state_shapes(cfg), 0, w_std=0.02)
    specs = block_specs(cfg)[0]
    clip = synth_input("bench.clip", (1, 3, 16, 224, 224), 1234)
    times = []
    t_begin = time.time()
    with torch.no_grad():
        O.svit_forward(clip, params, specs, cfg)  # warm-up
        while len(times) < max_runs and time.time() - t_begin < max_seconds:
            t0 = time.perf_counter()
            O.svit_forward(clip, params, specs, cfg)
            times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    times, cores = cpu_forward_timing(max_seconds=120.0, max_runs=steps + args.warmup)
    times = times[min(args.warmup, len(times) - 1):] or times
    sec = statistics.median(times)
    val = 1.0 / sec
    line = {"impl": "reference", "metric": "clips/sec (16x224^2, bf16) SViT forward", "value": val, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SViT ssv2.yaml forward, 16x224^2 clips, 4 object tokens/frame; each step = a "
                                   "batch-1 sample of the batch-64 workload on the host cores"},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"{len(times)} batch-1 forwards of the oracle port (torch CPU fp32)"},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="svit_b200", choices=["svit_b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel breakdown here")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--host-input", default="uint8", choices=["uint8", "bf16"],
                    help="e2e leg: what crosses PCIe each step -- decoded uint8 frames [B,T,H,W,3] (normalised on the "
                         "device by svit_normalize_u8, datasets/utils.py:287-303) or pre-normalised bf16 clips")
    ap.add_argument("--optimizer", action="store_true",
                    help="train mode: finish the step with the fused clip_grad_norm(1.0) + AdamW update "
                         "(svit_b200.optim, SOLVER settings of configs/ssv2.yaml)")
    ap.add_argument("--frames-pass", action="store_true",
                    help="train mode: also run the reference's no-grad pass over the B*16 single frames "
                         "(TRAIN.FORWARD_VIDEO_FRAMES, tools/train_net.py:105-110)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: configs[1] (the headline metric); train: configs[2], fwd + bwd + gradient all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)

    import torch.distributed as dist

    import svit_b200
    from svit_b200 import ops
    from svit_b200.config import ssv2_cfg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    B = args.batch
    cfg = ssv2_cfg()
    torch.manual_seed(0)
    model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).to(dev).eval()
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(B, 3, 16, 224, 224, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(2)]
    dev_in = [h.to(dev, non_blocking=True) for h in host]
    probs_host = torch.empty(B, cfg.MODEL.NUM_CLASSES, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    sampler = ClockSampler(local_rank)
    # The forward is captured once in a CUDA graph (svit_b200.GraphedForward) and replayed: one graph launch per step
    eager = model
    if not args.no_graph:
        model = svit_b200.GraphedForward(eager, dev_in[0])
    # ---- device-resident throughput
    with torch.no_grad():
        for i in range(W):
            model([dev_in[i & 1]] if args.no_graph else dev_in[i & 1])
        barrier()
        if rank == 0:
            sampler.start()
        l0 = ops.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            out, _ = model([dev_in[i & 1]] if args.no_graph else dev_in[i & 1])
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = ops.launches() - l0

        # ---- end to end: pinned host clips -> H2D (copy stream, double buffered) -> forward -> D2H of the probabilities
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()
        u8 = args.host_input == "uint8"
        if u8:  # frames as the decoder leaves them: one normalise kernel (svit_normalize_u8) in front of the same graph
            host = [torch.randint(0, 256, (B, 16, 224, 224, 3), generator=gen, dtype=torch.uint8).pin_memory()
                    for _ in range(2)]
        staged = [torch.empty(host[0].shape, dtype=host[0].dtype, device=dev) for _ in range(2)]
        mean, std = cfg.DATA.MEAN, cfg.DATA.STD

        def model_e2e(inp):
            x = inp[0] if args.no_graph else inp
            if u8:
                x = ops.normalize_u8(x, mean, std, torch.bfloat16)
            return model([x]) if args.no_graph else model(x)
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def e2e_loop(n):
            for i in range(n):
                j = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[j])
                    staged[j].copy_(host[j], non_blocking=True)
                    ready[j].record(copy_stream)
                main_stream.wait_event(ready[j])
                out, _ = model_e2e([staged[j]] if args.no_graph else staged[j])
                freed[j].record(main_stream)
                probs_host.copy_(out, non_blocking=True)

        for j in range(2):
            freed[j].record(main_stream)
        e2e_loop(2)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        copy_stream.wait_event(s0)
        e2e_loop(K)
        main_stream.wait_stream(copy_stream)
        s1.record()
        barrier()
        ms_e2e = max_over_ranks(s0.elapsed_time(s1))
        clocks = sampler.stop() if rank == 0 else None

        # ---- per-kernel breakdown (separate instrumented pass, CUDA events around every C-ABI call)
        prof = None
        if rank == 0:
            ops.profile_start()
            for i in range(2):
                eager([dev_in[i & 1]])
            torch.cuda.synchronize()
            prof = ops.profile_stop(steps=2)

    value = world * B * K / (ms / 1e3)
    e2e_val = world * B * K / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.barrier()  # stay alive until rank 0 has finished its instrumented pass: ordered NCCL teardown
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    wk = work_per_clip(cfg)
    fam = prof["families"]
    tot_ms = sum(f["ms_per_step"] for f in fam.values())
    tflops_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])

    def tensor_roof(name, flops_per_clip):
        t = fam.get(name, {}).get("ms_per_step", 0.0)
        ach = flops_per_clip * B / (t / 1e3) / 1e12 if t > 0 else 0.0
        return {"bound": "tensor", "achieved": ach, "peak": tflops_peak, "unit": "TFLOP/s", "frac": ach / tflops_peak,
                "traffic": None, "kernel": name, "share_of_step": t / tot_ms if tot_ms else None,
                "peak_source": f"{peaks['_source']} (sustained bf16 GEMM)"}

    roof_attn = tensor_roof("attention", wk["attn"])
    roof_gemm = tensor_roof("gemm", wk["gemm"])
    t_pool = fam.get("pool_ln", {}).get("ms_per_step", 0.0)
    ach_pool = wk["pool_bytes_bf16"] * B / (t_pool / 1e3) / 1e9 if t_pool > 0 else 0.0
    roof_pool = {"bound": "hbm", "achieved": ach_pool, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": ach_pool / peaks["hbm_gbs"], "traffic": None, "kernel": "pool_ln",
                 "share_of_step": t_pool / tot_ms if tot_ms else None, "peak_source": peaks["_source"]}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        tr = json.load(open(traffic_path))
        for r in (roof_attn, roof_gemm, roof_pool):
            r["traffic"] = tr.get(r["kernel"])
    # dominant KERNEL = the C-ABI call (one kernel shape) with the largest share of the step; its roofline is
    # algorithmic work per launch / average launch duration (CUDA events of the instrumented pass)
    dominant = max((roof_attn, roof_gemm, roof_pool), key=lambda r: r["share_of_step"] or 0.0)
    import re
    best = None
    for tag, d in prof["detail"].items():
        per_launch_ms = d["ms_per_step"] / max(d["calls_per_step"], 1e-9)
        m = re.match(r"svit_gemm\[(\d+)x(\d+)x(\d+)\]", tag)
        a = re.match(r"svit_attn_fwd\[B(\d+) h(\d+) Nq(\d+) Nk(\d+)\]", tag)
        if m:
            M_, N_, K_ = (int(x) for x in m.groups())
            work, bound = 2.0 * M_ * N_ * K_, "tensor"
        elif a:
            B_, h_, Nq_, Nk_ = (int(x) for x in a.groups())
            work, bound = 4.0 * B_ * h_ * Nq_ * Nk_ * 96, "tensor"
        else:
            continue
        if best is None or d["ms_per_step"] > best[1]["ms_per_step"]:
            best = (tag, d, work, bound, per_launch_ms)
    if best is not None:
        tag, d, work, bound, per_launch_ms = best
        ach = work / (per_launch_ms / 1e3) / 1e12
        tr = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
        dominant = {"bound": bound, "achieved": ach, "peak": tflops_peak, "unit": "TFLOP/s", "frac": ach / tflops_peak,
                    "traffic": tr.get(tag), "kernel": tag, "launches_per_step": d["calls_per_step"],
                    "avg_launch_ms": per_launch_ms, "share_of_step": d["ms_per_step"] / tot_ms if tot_ms else None,
                    "peak_source": f"{peaks['_source']} (sustained bf16 GEMM: the kernel runs inside a long step)",
                    "peak_burst": peaks.get("bf16_tflops")}

    line = {"metric": "clips/sec (16x224^2, bf16) SViT forward", "value": value, "unit": "clips/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"SViT (configs/ssv2.yaml: MViTv2-S 16x224^2 + 4 object tokens/frame) inference forward, "
                                   f"batch {B} clips per GPU, random init",
                       "parallelism": f"dp{world}", "global_batch": B * world,
                       "l2_policy": "inputs larger than L2 (308 MB of bf16 clips per step, >1 GB activations)",
                       "host_input_dtype": args.host_input, "launch": "eager" if args.no_graph else "cuda graph replay"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_val, "unit": "clips/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": host[0].numel() * host[0].element_size(),
                    "d2h_bytes_per_step": probs_host.numel() * probs_host.element_size()},
            "roofline": dominant, "roofline_attention": roof_attn, "roofline_gemm": roof_gemm,
            "roofline_pool_ln": roof_pool,
            "kernel_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(fam.items())}}
    if world == 1 and not args.no_cpu_baseline:
        times, cores = cpu_forward_timing(max_seconds=20.0, max_runs=5)
        sec = statistics.median(times)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": f"{len(times)} batch-1 fp32 forwards of the CPU oracle port "
                                          f"(1 of the {B} clips of a step), median {sec:.3f} s"}
    if args.profile_json:
        with open(args.profile_json, "w") as f:
            json.dump(prof, f, indent=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """configs[2]: full training step (forward + backward, bf16 activations, fp32 parameter gradients) with batch 8
    clips per GPU sharded over the ranks and one bucketed NCCL gradient all-reduce per step (no optimizer: the
    optimizer is outside the hot path)."""
    import torch.distributed as dist

    import svit_b200
    from svit_b200 import ops
    from svit_b200.config import ssv2_cfg
    from svit_b200.distributed import GradAllReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(3, args.warmup), args.steps
    B = 8 if args.batch == 64 else args.batch
    cfg = ssv2_cfg()
    torch.manual_seed(0)
    model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).to(dev).train()
    reducer = GradAllReducer(model.parameters()) if world > 1 else None
    optimizer = None
    if args.optimizer:
        from svit_b200.optim import construct_optimizer
        optimizer = construct_optimizer(model, cfg)
    gen = torch.Generator().manual_seed(1234 + rank)
    clips = [torch.randn(B, 3, 16, 224, 224, generator=gen).to(torch.bfloat16).to(dev) for _ in range(2)]
    labels = torch.randint(0, cfg.MODEL.NUM_CLASSES, (B,), generator=gen).to(dev)

    def step(i):
        for p in model.parameters():
            p.grad = None
        if reducer is not None:
            reducer.prepare()
        preds, extra = model([clips[i & 1]])
        loss = torch.nn.functional.cross_entropy(extra["logits"].float(), labels)
        if args.frames_pass:
            from svit_b200.distributed import consistency_loss, forward_video_frames
            _p, _e = forward_video_frames(model, clips[i & 1])
            for k, v in consistency_loss(model._lambda, extra, _e).items():  # empty with the stock lambda keys
                loss = loss + model._lambda[k] * v
        loss.backward()
        if reducer is not None:
            reducer.finish()
        if optimizer is not None:
            optimizer.step(max_norm=cfg.SOLVER.CLIP_GRAD_L2NORM)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        line = {"metric": "clips/sec (16x224^2, bf16) SViT training step (fwd+bwd+grad all-reduce)",
                "value": world * B * K / (ms / 1e3), "unit": "clips/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "gpu_launches": ops.launches() - l0, "loss": float(loss),
                "config": {"workload": f"SViT (configs/ssv2.yaml) training step, batch {B} clips per GPU, random init, "
                                       "cross-entropy on the class logits, no optimizer step"
                                       .replace("no optimizer step", "fused clip_grad_norm + AdamW step" if args.optimizer
                                                else "no optimizer step")
                                       + (", + no-grad frames pass (B*16 frames, T=1)" if args.frames_pass else ""),
                           "parallelism": f"dp{world}", "global_batch": B * world}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
