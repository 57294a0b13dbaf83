#!/usr/bin/env python
"""Benchmark of the reverse-diffusion hot path (BASELINE.json metric: mel frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl reference]

A "step" is one full reverse diffusion (all K_diff Denoiser calls + posterior updates + denorm/mask)
over one synthetic batch.  Workload at any N: BASELINE configs[1] — LJSpeech `naive`, K_diff=4,
B=64 utterances x T=800 frames per GPU (weak scaling: utterances shard across ranks with no
collective on the data path).  One JSON line is printed by rank 0.

  value     mel frames/s with inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e       the same metric through the public module API with HOST (pinned) inputs: H2D copy of
            cond + mask, on-device noise draw, D2H read of the mel, all inside the timed region
  roofline  the dominant kernel against the measured tensor peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle port timed on this box's host cores on a bounded sample (rank 0, N=1)

`--impl reference` times the reference's CPU implementation of the path (the torch-CPU oracle port;
the reference itself is Python and cannot travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_FRAME_STEP = 23_805_952          # SURVEY.md §8d / BASELINE.md §3 (algorithmic, no halo)
CONV_FLOPS_PER_FRAME_STEP = 786_432        # the k=3 conv alone (dominant kernel of the fp32 path)
B_PER_GPU, T_FRAMES, K_DIFF = 64, 800, 4   # BASELINE configs[1]
CPU_SAMPLE_B = 16                          # BASELINE configs[0] shape for the CPU legs
METRIC, UNIT = "mel_frames_per_sec", "frames/s"


def ncu_traffic(B, T):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (same workload), or None."""
    import glob
    if (B, T) != (B_PER_GPU, T_FRAMES):
        return None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "ncu_traffic.json")), reverse=True):
        try:
            return float(json.load(open(p))["mean_bytes_per_launch"])
        except Exception:
            continue
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_tflops": float(d["bf16_tflops_sustained"]), "hbm_gbs": float(d["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"tensor_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        pmax = max(pw)
        busy = [c for c, w in zip(sm, pw) if w >= 0.6 * pmax] or sm      # samples taken under load
        return {"sm_mhz": statistics.median(busy), "sm_min_mhz": min(busy), "sm_max_mhz": max(mx), "power_w_max": pmax,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(busy)}


# ---------------------------------------------------------------------------------------------
def cpu_oracle_throughput(repeats: int, B: int = CPU_SAMPLE_B):
    """Frames/s of the CPU oracle port on all host threads, best of `repeats` after one warm-up."""
    import torch
    from mixgan_tts_b200 import configs, synth
    from oracle.diffusion import DiffusionOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, _, mc, _ = configs.make_configs("LJSpeech", "naive")
    W = synth.make_denoiser_weights(0)
    orc = DiffusionOracle(W, model="naive", denoiser_cfg=mc["denoiser"], spec_min=[configs.SPEC_MIN] * 80,
                          spec_max=[configs.SPEC_MAX] * 80)
    inp = synth.make_inputs(1234, B, T_FRAMES, K_DIFF)
    tt = lambda k: torch.from_numpy(inp[k])
    args = (tt("cond"), None, tt("pad_mask"))
    kw = dict(x_T=tt("x_T"), noises=tt("noises"))
    small = synth.make_inputs(1, 1, 64, K_DIFF)
    orc.forward_inference(torch.from_numpy(small["cond"]), None, torch.from_numpy(small["pad_mask"]),
                          x_T=torch.from_numpy(small["x_T"]), noises=torch.from_numpy(small["noises"]))  # warm-up
    times = []
    for _ in range(max(1, repeats)):
        t0 = time.perf_counter()
        orc.forward_inference(*args, **kw)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": B * T_FRAMES / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"LJSpeech naive K=4, B={B} x T={T_FRAMES} (BASELINE configs[0] shape), fp32 torch-CPU oracle, "
                      f"best of {len(times)} after warm-up, {best:.2f} s per pass, torch {torch.__version__}"}, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train":
        base = cpu_oracle_train_throughput(args.train_batch)
        print(json.dumps({
            "impl": "reference", "metric": "train_frames_per_sec", "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": 2, "warmup": 0, "ms_per_step": 1e3 * args.train_batch * T_FRAMES / base["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LJSpeech naive diffusion-decoder training branch (forward + autograd backward), CPU, "
                                   f"B={args.train_batch} x T={T_FRAMES}"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    steps = max(1, min(args.steps, 5))
    base, times = cpu_oracle_throughput(steps)
    ms = 1e3 * statistics.mean(times)
    val = CPU_SAMPLE_B * T_FRAMES / statistics.mean(times)
    base = dict(base, value=val)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LJSpeech naive K=4 reverse diffusion, CPU, bounded sample B={CPU_SAMPLE_B} x T={T_FRAMES} "
                               "per step of the BASELINE configs[1] workload"},
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from mixgan_tts_b200 import GaussianDiffusion, _lib, configs, synth
    from mixgan_tts_b200.modules import PRECISIONS
    lib = _lib.load()
    dims0 = _lib.ModelDims(80, 256, 256, 20, 0)
    prec = args.precision
    if prec == "auto":
        prec = "bf16" if lib.mgb_packed_bytes(C.byref(dims0), _lib.PREC_BF16) > 0 else "fp32"

    cfg = configs.make_configs("LJSpeech", "naive")
    gd = GaussianDiffusion(*cfg, precision=prec)
    W = synth.make_denoiser_weights(0)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
    gd = gd.to(dev).eval()
    B, T, K = args.batch, T_FRAMES, gd.num_timesteps
    den = gd.denoise_fn

    # Rotating input sets so that no step finds its inputs in L2 (3 x ~134 MB > 126 MB L2).
    NSETS = 3
    sets = []
    for i in range(NSETS):
        inp = synth.make_inputs(1234 + 17 * i + 1000 * rank, B, T, K)
        sets.append({"cond": torch.from_numpy(inp["cond"]).to(dev), "pad": torch.from_numpy(inp["pad_mask"]).to(dev),
                     "x_T": torch.from_numpy(inp["x_T"]).to(dev), "noises": torch.from_numpy(inp["noises"]).to(dev)})
    set_bytes = sum(v.numel() * v.element_size() for v in sets[0].values())
    pad_u8 = [s["pad"].to(torch.uint8).contiguous() for s in sets]

    def step_resident(i):
        s = sets[i % NSETS]
        return gd._sample_core(s["cond"], None, s["x_T"], s["noises"], pad_u8[i % NSETS], want_states=False)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident timing
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.mgb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_resident(i)
    ev1.record()
    barrier()
    launches = lib.mgb_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)

    # dominant-kernel duration: a short separate pass in which the library brackets every launch of the fused
    # kernel with CUDA events on the launching stream (kept out of the pass above: the events cost a few us each)
    prof_steps = min(args.steps, 4)
    lib.mgb_profile_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(prof_steps):
        step_resident(i)
    p1.record()
    torch.cuda.synchronize()
    ms_prof_pass = p0.elapsed_time(p1)
    ktot, kcnt = C.c_float(0), C.c_int(0)
    _lib.check(lib.mgb_profile_collect(C.byref(ktot), C.byref(kcnt)), "mgb_profile_collect")
    lib.mgb_profile_enable(0)

    # ------------------------------------------------------------------ end-to-end timing (host buffers)
    # The public batch-synthesis harness: pinned host cond/mask in, pinned host mel out; the H2D copy of batch i+1
    # and the D2H copy of batch i-1 overlap the reverse diffusion of batch i.  Every step's copies are inside the
    # timed region; the noise is drawn on the device inside GaussianDiffusion.forward, as the reference does.
    from mixgan_tts_b200.pipeline import BatchSynthesizer
    host = [(s["cond"].cpu().pin_memory(), s["pad"].cpu().pin_memory()) for s in sets]
    synth_pipe = BatchSynthesizer(gd, dev)

    def run_e2e(n):
        acc = 0.0
        for mel in synth_pipe.run(host[i % NSETS] for i in range(n)):
            acc += float(mel[0, 0, 0])          # touch the host result of every step
        return acc

    run_e2e(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    from mixgan_tts_b200 import shard
    ms_total = shard.max_over_ranks(ms_total, dev)
    ms_e2e = shard.max_over_ranks(ms_e2e, dev)

    if rank == 0:
        pk = peaks()
        frames_per_step = world * B * T
        value = frames_per_step * args.steps / (ms_total * 1e-3)
        e2e_val = frames_per_step * args.steps / (ms_e2e * 1e-3)
        k_ms = ktot.value / max(kcnt.value, 1)
        calls = prof_steps * K                      # Denoiser calls in the profiled pass
        if prec in ("bf16", "fp16"):
            # one Denoiser call = `launches_per_call` launches of fused_group_kernel (layer groups);
            # algorithmic FLOPs per launch = FLOPs per call / launches per call
            per_call = max(kcnt.value // calls, 1)
            kern = f"fused_pair_kernel (tcgen05 cta_group::2; {per_call} launches = one Denoiser call over the batch)"
            fl = FLOPS_PER_FRAME_STEP / per_call
        else:
            kern, fl = "conv_gemm_kernel<EPI_GATE> (k=3 conv + gate of one block, fp32 CUDA cores)", CONV_FLOPS_PER_FRAME_STEP
        achieved = fl * B * T / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        valid = float(sum(int((~s["pad"]).sum()) for s in sets)) / NSETS
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16"}.get(prec, "f32"), "data": "synthetic",
            "config": {"workload": f"LJSpeech naive K={K} reverse diffusion, B={B} x T={T} per GPU (BASELINE configs[1]), "
                                   "random-init Denoiser, fixed injected noise",
                       "precision": prec, "l2": f"inputs rotate over {NSETS} sets of {set_bytes / 1e6:.0f} MB (> 126 MB L2)",
                       "frames_valid_per_s": value * valid / (B * T), "frame_steps_per_s": value * K,
                       "rtf_valid_audio": (ms_total / args.steps * 1e-3) / (world * valid * 256 / 22050)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tensor_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tensor_tflops"],
                         "traffic": ncu_traffic(B, T) if prec == "bf16" else None, "kernel": kern,
                         "kernel_ms": k_ms, "kernel_launches_timed": kcnt.value,
                         "kernel_share_of_step": ktot.value / ms_prof_pass if ms_prof_pass > 0 else None,
                         "whole_step_tflops": FLOPS_PER_FRAME_STEP * K * B * T * args.steps / (ms_total * 1e-3) / 1e12,
                         "peak_source": pk["source"]},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(B * T * 256 * 4 + B * T),
                    "d2h_bytes_per_step": int(B * T * 80 * 4), "ms_per_step": ms_e2e / args.steps,
                    "api": "BatchSynthesizer.run (GaussianDiffusion.forward per batch; pinned host cond/mask in, pinned host mel "
                           "out; copies of neighbouring batches overlap compute on separate streams)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_oracle_throughput(8)      # ~10 s of CPU work on the box's host cores
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def run_train(args):
    """--workload train: BASELINE configs[4] restricted to this repo's path — the diffusion decoder's training branch
    (GaussianDiffusion.forward with mel given: q_sample x2, Denoiser forward, clamp, posterior sample), backward through
    the library's Denoiser backward, data-parallel gradient all-reduce (NCCL, bucketed, overlapped with the backward) and
    a fused Adam step on the Denoiser's parameters.  Per GPU: B=8 utterances x T=800 frames (config/LJSpeech/train.yaml:6).
    The JCU discriminator / FastSpeech2 encoder of the full training step stay the reference's torch code (out of scope)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from mixgan_tts_b200 import GaussianDiffusion, _lib, configs, synth, shard
    from mixgan_tts_b200.grad_sync import GradSync
    lib = _lib.load()
    B, T = args.train_batch, T_FRAMES
    cfg = configs.make_configs("LJSpeech", "naive")
    prec = "bf16" if args.precision == "auto" else args.precision
    gd = GaussianDiffusion(*cfg, precision=prec)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_denoiser_weights(0).items()})
    gd = gd.to(dev).train()
    K = gd.num_timesteps
    opt = torch.optim.Adam(gd.denoise_fn.parameters(), lr=1e-5, fused=True)
    sync = GradSync() if world > 1 else None
    NSETS = 3
    sets = []
    for i in range(NSETS):
        inp = synth.make_inputs(77 + 13 * i + 1000 * rank, B, T, K)
        ex = synth.make_train_extras(78 + 13 * i + 1000 * rank, B, T, K)
        pr = synth.grad_probe(79 + 13 * i + 1000 * rank, B, T)
        to = lambda a: torch.from_numpy(a).to(dev)
        sets.append({"cond": to(inp["cond"]), "pad": to(inp["pad_mask"]), "mel": to(ex["mel"]), "r0": to(pr["r0"]), "r1": to(pr["r1"])})

    def step(i, with_sync=True):
        s = sets[i % NSETS]
        gd.denoise_fn.grad_sync = sync if with_sync else None
        opt.zero_grad(set_to_none=True)
        cond = s["cond"].detach().requires_grad_(True)     # the encoder would receive d loss / d cond
        out = gd(s["mel"], cond, None, s["pad"])
        loss = (out[0] * s["r0"]).sum() + (out[3] * s["r1"]).sum()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, **kw):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(i, **kw)
        e1.record()
        barrier()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev)

    for i in range(max(args.warmup, 3)):
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.mgb_launch_count()
    ms = timed(args.steps)
    launches = lib.mgb_launch_count() - n0
    ms_nosync = timed(args.steps, with_sync=False) if world > 1 else ms
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        frames = world * B * T
        flops = 3 * FLOPS_PER_FRAME_STEP * frames            # forward + data-grad + weight-grad GEMMs
        line = {"metric": "train_frames_per_sec", "value": frames * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec == "bf16" else "f32", "data": "synthetic",
                "config": {"precision": prec, "workload": f"LJSpeech naive diffusion-decoder training branch: Denoiser fwd+bwd + fused Adam, B={B} x T={T} "
                                       "per GPU (BASELINE configs[4] restricted to the Denoiser path), gradient all-reduce over NCCL",
                           "tflops": flops * args.steps / (ms * 1e-3) / 1e12,
                           "allreduce_exposed_ms_per_step": (ms - ms_nosync) / args.steps,
                           "grad_bytes": int(lib.mgb_flat_weight_count(C.byref(gd.denoise_fn.dims))) * 4},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_train_throughput(B)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c3(args):
    """--workload c3: BASELINE configs[2] / SURVEY 8(d) C3 — LJSpeech `shallow` (aux-decoder mel -> K=1 shallow diffusion),
    512 utterances of T=800 sharded contiguously by utterance over the N ranks (STRONG scaling, no collective on the data
    path), every rank running mixgan_tts_b200.pipeline.BatchSynthesizer on its shard in batches of 64 with pinned HOST
    buffers in and out.  value = 512 * 800 frames / (max over ranks of the device-timed pass)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from mixgan_tts_b200 import GaussianDiffusion, _lib, configs, synth, shard
    from mixgan_tts_b200.pipeline import BatchSynthesizer
    lib = _lib.load()
    N_UTT, T, BATCH = 512, T_FRAMES, 64
    cfg = configs.make_configs("LJSpeech", "shallow")
    gd = GaussianDiffusion(*cfg, precision="bf16" if args.precision == "auto" else args.precision)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_denoiser_weights(0).items()})
    gd = gd.to(dev).eval()
    lo, hi = shard.contiguous_shard(N_UTT, world, rank)
    # the shard's utterances as pinned host batches (3 distinct synthetic batches, cycled: > L2 per pass)
    protos = []
    for i in range(3):
        inp = synth.make_inputs(500 + 7 * i + 100 * rank, BATCH, T, 1, shallow=True)
        protos.append(tuple(torch.from_numpy(inp[k]).pin_memory() for k in ("cond", "pad_mask")) + (None,)
                      + (torch.from_numpy(inp["coarse_mel"]).pin_memory(),))
    batches = []
    for b0 in range(lo, hi, BATCH):
        n = min(BATCH, hi - b0)
        pr = protos[(b0 // BATCH) % 3]
        batches.append(tuple(None if t is None else t[:n] for t in pr))
    pipe = BatchSynthesizer(gd, dev)
    dev_batches = [tuple(None if t is None else t.to(dev) for t in b) for b in batches]

    def pass_e2e():
        acc = 0.0
        for mel in pipe.run(iter(batches)):
            acc += float(mel[0, 0, 0])
        return acc

    def pass_resident():
        for cond, pad, spk, coarse in dev_batches:
            gd(None, cond, spk, pad, coarse_mel=coarse)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev)

    n0 = lib.mgb_launch_count()
    ms = timed(pass_resident)
    launches = lib.mgb_launch_count() - n0
    ms_e2e = timed(pass_e2e)
    if rank == 0:
        val = N_UTT * T * args.steps / (ms * 1e-3)
        print(json.dumps({
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"LJSpeech shallow K=1 batch synthesis, {N_UTT} utterances x T={T} sharded by utterance over "
                                   f"{world} GPU(s) in batches of {BATCH} (BASELINE configs[2]); value: inputs resident in HBM "
                                   "(shallow start + K=1 reverse diffusion + denorm per batch, noise drawn on the device)",
                       "utterances_per_rank": hi - lo},
            "e2e": {"value": N_UTT * T * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(N_UTT * T * (256 * 4 + 80 * 4 + 1)), "d2h_bytes_per_step": int(N_UTT * T * 80 * 4),
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "BatchSynthesizer.run on the rank's shard: pinned host cond / mask / coarse mel in, pinned host mel out; "
                           "the H2D copy of a 64-utterance batch (69 MB) takes longer than its compute, so this arm is PCIe-bound"},
            "gpu_launches": int(launches)}))
    if world > 1:
        dist.destroy_process_group()


def cpu_oracle_train_throughput(B: int):
    """The oracle's training branch + torch autograd on all host threads, one step at the same shape."""
    import torch
    from mixgan_tts_b200 import configs, synth
    from oracle.diffusion import DiffusionOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, _, mc, _ = configs.make_configs("LJSpeech", "naive")
    W = synth.make_denoiser_weights(0)
    orc = DiffusionOracle(W, model="naive", denoiser_cfg=mc["denoiser"], spec_min=[configs.SPEC_MIN] * 80, spec_max=[configs.SPEC_MAX] * 80)
    inp, ex, pr = synth.make_inputs(77, B, T_FRAMES, K_DIFF), synth.make_train_extras(78, B, T_FRAMES, K_DIFF), synth.grad_probe(79, B, T_FRAMES)
    tt = lambda a: torch.from_numpy(a)
    Wt = {k: tt(v).requires_grad_(True) for k, v in W.items()}
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        cond = tt(inp["cond"]).requires_grad_(True)
        out = orc.forward_training_graph(tt(ex["mel"]), cond, None, tt(inp["pad_mask"]), t=tt(ex["t"]).clone(), noise_t=tt(ex["noise_t"]),
                                         noise_prev=tt(ex["noise_prev"]), post_noise=tt(ex["post_noise"]), W=Wt)
        ((out[0] * tt(pr["r0"])).sum() + (out[3] * tt(pr["r1"])).sum()).backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": B * T_FRAMES / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one training-branch step (forward + autograd backward, no optimizer) at B={B} x T={T_FRAMES}, fp32 torch-CPU oracle, "
                      f"best of 2, {best:.2f} s, torch {torch.__version__}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "c3"],
                    help="sample = the headline reverse-diffusion benchmark; train = Denoiser training step (configs[4] path); "
                         "c3 = 512-utterance shallow-diffusion batch synthesis, strong scaling (configs[2])")
    ap.add_argument("--train-batch", type=int, default=8)
    args = ap.parse_args()
    if args.workload == "train" and args.impl == "ours":
        run_train(args)
    elif args.workload == "c3" and args.impl == "ours":
        run_c3(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
