#!/usr/bin/env python
"""Benchmark of the speaker-embedding hot path (BASELINE.json metric: speaker embeddings/s on 10 s clips).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode 0|1]

A "step" is one pass of VoiceEncoder + CAMPPlus over one batch of 256 synthetic ten-second 16 kHz clips per GPU
(BASELINE.json configs[1]).  `value` = whole-job clips/s with the PCM already resident in HBM; `e2e` = the same through
cbx_embed_host with HOST buffers (host->device copy of the PCM and device->host read of the embeddings inside the timed
region).  N>1 (torchrun): weak scaling, every rank embeds its own batch, then ONE all-gather of the (256, 448) block.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS = 256
CLIP_SAMPLES = 160000
WORKLOAD = "256 x 10 s 16 kHz clips per GPU, VoiceEncoder(256-d)+CAMPPlus(192-d), random-init weights (BASELINE configs[1])"
# algorithmic FLOPs per 10 s clip
FLOPS_PER_CLIP = 5190451200 + 1572864 + 11234711552 + 338017680 + 451415360


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons.  The sampler is started EARLY (nvidia-smi needs a few hundred ms to deliver its
    first line) and every sample carries a timestamp; `summary(t0, t1)` keeps the samples that fall inside the timed region
    (wall-clock window), widening to the nearest ones only if the region was shorter than the sampling period."""

    def __init__(self, index: int, period_ms: int = 20):
        self.index, self.proc, self.samples, self.period_ms = index, None, [], period_ms

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.samples.append((time.time(), ln))
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=2)

    def summary(self, t0: float, t1: float):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [(t, ln) for t, ln in self.samples if t0 <= t <= t1 + self.period_ms / 1e3]
        how = "inside the timed region"
        if not inside and self.samples:      # region shorter than a sampling period: the samples bracketing it
            inside = sorted(self.samples, key=lambda s: min(abs(s[0] - t0), abs(s[0] - t1)))[:2]
            how = "nearest to the timed region"
        sm, mx, reasons = [], 0.0, set()
        for _, ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "sampled": how}


def make_batch(rank: int):
    from chatterbox_embed_b200 import synth
    wavs = [synth.clip(rank * CLIPS + i, CLIP_SAMPLES) for i in range(CLIPS)]
    off = np.arange(CLIPS + 1, dtype=np.int64) * CLIP_SAMPLES
    return wavs, off


def cpu_baseline(n_clips: int, threads: int):
    """Oracle port of the reference path, timed on the host cores (B=1 loop, the reference's real usage)."""
    import torch
    from chatterbox_embed_b200 import synth
    from oracle import nets, weights
    torch.set_num_threads(threads)
    sdv, sdc = weights.ve_state_dict("W0"), weights.campplus_state_dict("W0")
    wavs = [synth.clip(i, CLIP_SAMPLES) for i in range(n_clips)]
    nets.ve_embed_wavs(sdv, wavs[:1]); nets.campplus_embed_wavs(sdc, wavs[:1])      # warm-up
    t0 = time.perf_counter()
    nets.ve_embed_wavs(sdv, wavs)
    nets.campplus_embed_wavs(sdc, wavs)
    dt = time.perf_counter() - t0
    return n_clips / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 32
    for _ in range(max(args.warmup, 1) - 1):
        cpu_baseline(2, threads)
    vals, t_total = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_baseline(per_step, threads)
        vals.append(v); t_total += dt
    value = per_step * args.steps / t_total
    line = {"impl": "reference", "metric": "speaker embeddings/sec (10 s clips)", "value": value, "unit": "clips/s",
            "audio_s_per_s": value * 10.0, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port",
                             "sample": f"{per_step} of the 256 ten-second clips per step, B=1 loop, oracle port of the reference "
                                       "(torch CPU fp32); the verbatim reference cannot travel to the GPU box"},
            "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", type=int, default=int(os.environ.get("CBX_MODE", "1")))
    ap.add_argument("--opt", action="append", default=[], help="libcbx option key=value (cbx_set_option)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    clk = ClockSampler(local).start()          # early: the first nvidia-smi line takes a few hundred ms

    torch.manual_seed(0)                       # random-init weights of the same architecture (no checkpoints offline)
    ve = VoiceEncoder().to(dev).eval()
    cp = CAMPPlus().to(dev).eval()
    emb = scheduler.SpeakerEmbedder(ve, cp)
    ctx = emb.ctx()
    ctx.set_option("mode", args.mode)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))

    wavs, off = make_batch(rank)
    host = torch.empty(CLIPS * CLIP_SAMPLES, dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    for i, w in enumerate(wavs):
        host_np[off[i]:off[i + 1]] = w
    pcm = host.to(dev)
    shards = [np.arange(CLIPS) + r * CLIPS for r in range(world)]
    gathered = torch.empty((world * CLIPS, scheduler.EMB), dtype=torch.float32, device=dev) if world > 1 else None

    def step_device():
        ve_o, xv_o, status = emb.embed_device(pcm, off)
        if world > 1:
            block = torch.cat([ve_o, xv_o], dim=1)
            dist.all_gather_into_tensor(gathered, block)
        return ve_o, xv_o

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    # ---- device-resident throughput -----------------------------------------------------------------------------
    for _ in range(W):
        step_device()
    sync_all()
    l0 = ctx.launch_count()
    tw0 = time.time()
    ms, _ = timed(step_device, K)
    tw1 = time.time()
    launches = ctx.launch_count() - l0
    clocks = clk.summary(tw0, tw1)
    value = world * CLIPS * K / (ms / 1e3)

    clk.stop()

    # ---- end to end through host buffers (cbx_embed_host) -------------------------------------------------------
    flags_pinned = _lib.DO_VE | _lib.DO_XV | _lib.PCM_PINNED
    out_holder = {}

    def step_host():
        ve_o, xv_o, status = ctx.embed_host(host_np, off, 20.0, 77, 0.8, flags_pinned)
        out_holder["ve"], out_holder["xv"] = ve_o, xv_o

    for _ in range(2):
        step_host()
    _, wall_sync = timed(step_host, K)
    # the voice-bank call a user makes for many batches: SpeakerEmbedder.embed_stream, two batches in flight.  Every step
    # still copies its own 164 MB of PCM host->device and its embeddings device->host inside the timed region; the copies
    # of batch k+1 overlap the kernels of batch k.
    def steps_stream(k):
        n_done = 0
        for ve_o, xv_o, status in emb.embed_stream(((host_np, off) for _ in range(k)), pinned=True):
            out_holder["ve"], out_holder["xv"] = ve_o, xv_o
            n_done += 1
        assert n_done == k
    steps_stream(2)
    _, wall = timed(lambda: steps_stream(K), 1)
    e2e_value = world * CLIPS * K / wall
    e2e_sync_value = world * CLIPS * K / wall_sync

    # ---- per-kernel device times (CUDA events on the launching stream), separate profiled steps ------------------
    # (the two encoder chains are serialised for this pass so that a kernel's events time that kernel alone)
    ctx.set_option("overlap", 0)
    ctx.profile_enable(True)
    prof_steps = min(K, 2)
    for _ in range(prof_steps):
        step_device()
    torch.cuda.synchronize()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    ctx.set_option("overlap", 1)
    # kernel families: the per-conv tags of the FCM head ("fcm_conv_gemm:l1b0c1" ...) are one kernel
    fam = {}
    for k, v in prof.items():
        f = fam.setdefault(k.split(":")[0], dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
        for key in ("ms", "launches", "flops", "bytes"):
            f[key] += v[key]
    tot_ms = sum(v["ms"] for v in fam.values()) or 1.0
    tname, t = max(fam.items(), key=lambda kv: kv[1]["ms"])
    pk = peaks()
    tensor_peak = pk["bf16_sus"] / 2.0         # TF32 dense = half the bf16 rate; the bf16 sustained peak is measured
    ridge = tensor_peak * 1e12 / (pk["hbm"] * 1e9)          # FLOP per byte above which a kernel is tensor bound
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram__bytes_read+write per launch from ncu --set full captures
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(tname, {}).get("dram_bytes_per_launch")
    secs = t["ms"] / 1e3
    common = {"kernel": tname, "share_of_step": t["ms"] / tot_ms, "avg_launch_ms": t["ms"] / t["launches"],
              "launches_per_step": t["launches"] / prof_steps, "traffic": traffic,
              "arithmetic_intensity_flop_per_byte": (t["flops"] / t["bytes"]) if t["bytes"] else None}
    if t["bytes"] > 0 and (t["flops"] == 0 or t["flops"] / t["bytes"] < ridge):
        achieved = t["bytes"] / secs / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                "peak_source": f"{pk['src']} HBM copy bandwidth", "tflops": t["flops"] / secs / 1e12, **common}
    elif t["flops"] > 0:
        achieved = t["flops"] / secs / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "peak_source": f"{pk['src']} bf16 sustained / 2 (TF32 runs at half the bf16 rate; no TF32 peak is measured)", **common}
    else:
        roof = {"bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, **common}
    kernels = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                   "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] > 0 and v["ms"] > 0 else None,
                   "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["bytes"] > 0 and v["ms"] > 0 else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_s = 192                          # ~15-20 s of CPU work on the box's host cores
            v, dt = cpu_baseline(n_s, threads)
            cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                   "sample": f"{n_s} of the {CLIPS} ten-second clips, B=1 loop, {dt:.1f} s of CPU work (oracle port, torch CPU fp32)"}
        cat_bf16 = int(ctx.get_option("cat_bf16"))                # opt-in (--opt cat_bf16=1|2): not the parity mode, say so in the line
        line = {
            "metric": "speaker embeddings/sec (10 s clips)", "value": value, "unit": "clips/s",
            "audio_s_per_s": value * 10.0, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.mode == 0 else ("tf32" if not cat_bf16 else "tf32+bf16"), "data": "synthetic",
            "precision_note": None if args.mode == 0 else "tcgen05 kind::tf32 with fp32 accumulation everywhere; the front-end DFT is 3xTF32 "
                                                          "(hi/lo split, fp32-accurate); activations and state are stored in fp32"
                                                          + ("" if not cat_bf16 else f"; EXCEPT option cat_bf16={cat_bf16}: the D-TDNN bottleneck / transit GEMMs read a bf16 "
                                                             "copy of the concatenation buffers" + (" and run on bf16 operands (kind::f16)" if cat_bf16 == 2 else "")
                                                             + " -- a looser-tolerance setting, not the fp32/TF32 parity mode"),
            "config": {"workload": WORKLOAD, "mode": "strict-fp32 SIMT" if args.mode == 0 else "tcgen05 TF32",
                       "l2": "inputs (164 MB PCM per step) and activations exceed the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world}", "clips_per_gpu": CLIPS},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": CLIPS * CLIP_SAMPLES * 4,
                    "d2h_bytes_per_step": CLIPS * (256 + 192 + 1) * 4,
                    "api": "SpeakerEmbedder.embed_stream (cbx_embed_host_submit/_wait, two batches in flight)",
                    "single_call_value": e2e_sync_value, "single_call_api": "cbx_embed_host (copy, compute, copy back, sync)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "algorithmic_tflops": value * FLOPS_PER_CLIP / 1e12, "kernels": kernels,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
