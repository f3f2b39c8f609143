#!/usr/bin/env python
"""Benchmark of the speaker-embedding hot path (BASELINE.json metric: speaker embeddings/s on 10 s clips).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode 0|1] [--config 2|3|4|5] [--job-clips N] [--ragged]

--config 2 (default; BASELINE.json configs[1]): a "step" is one pass of VoiceEncoder + CAMPPlus over one batch of 256
synthetic ten-second 16 kHz clips per GPU.  --config 3 (configs[2]): one ragged batch of 1024 clips, 3-30 s each, per GPU.
`value` = whole-job clips/s with the PCM already resident in HBM; `e2e` = the same through SpeakerEmbedder.embed_stream with
HOST buffers (host->device copy of the PCM and device->host read of the embeddings inside the timed region; at N>1 the
all-gather of the embeddings is inside too).  After the timed loops a sample of the step's clips is compared with the CPU
oracle and reported as `parity` (the numbers the timed configuration itself produced, not a separate small test).
N>1 (torchrun): weak scaling, every rank embeds its own batch, then ONE all-gather of the (clips, 448) block.

--config 4 (configs[3], the voice-bank job): --job-clips N (default 100 000) clips are partitioned over the ranks by the
native scheduler (cbx_partition), every rank streams its shard through embed_stream in 256-clip batches from pinned host
memory, ONE NCCL all-gather returns all embeddings to every rank, they are un-permuted into clip order and a sample is
checked against the oracle on rank 0.  Strong scaling: `value` = N clips / job time (partition + stream + gather + un-permute).
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
WORKLOADS = {
    2: "256 x 10 s 16 kHz clips per GPU, VoiceEncoder(256-d)+CAMPPlus(192-d), random-init weights (BASELINE configs[1])",
    3: "ragged batch of 1024 clips, 3-30 s each (17 093 s of audio) per GPU, VoiceEncoder(256-d)+CAMPPlus(192-d), random-init weights (BASELINE configs[2])",
    5: "prepare_conditionals over 256 x 10 s 24 kHz prompts on one GPU: resample to 16 kHz, 24 kHz prompt mel, S3 tokenizer log-mel front end, VoiceEncoder + CAMPPlus, generator-side projections (BASELINE configs[4])",
    4: "voice-bank job: {n} clips ({kind}) sharded over the GPUs by cbx_partition, 256-clip batches streamed from pinned host memory, one NCCL all-gather (BASELINE configs[3])",
}
WORKLOAD = WORKLOADS[2]
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


def make_batch(rank: int, config: int = 2):
    from chatterbox_embed_b200 import synth
    if config == 3:
        lens = [int(x) for x in synth.ragged_lengths(1024, seed=2024 + rank)]
        wavs = [synth.clip(rank * 1024 + i, n) for i, n in enumerate(lens)]
    else:
        wavs = [synth.clip(rank * CLIPS + i, CLIP_SAMPLES) for i in range(CLIPS)]
    off = np.concatenate([[0], np.cumsum([len(w) for w in wavs])]).astype(np.int64)
    return wavs, off


def cpu_baseline(n_clips: int, threads: int, wavs=None):
    """Oracle port of the reference path, timed on the host cores (B=1 loop, the reference's real usage)."""
    import torch
    from chatterbox_embed_b200 import synth
    from oracle import nets, weights
    torch.set_num_threads(threads)
    sdv, sdc = weights.ve_state_dict("W0"), weights.campplus_state_dict("W0")
    if wavs is None:
        wavs = [synth.clip(i, CLIP_SAMPLES) for i in range(n_clips)]
    nets.ve_embed_wavs(sdv, wavs[:1]); nets.campplus_embed_wavs(sdc, wavs[:1])      # warm-up
    t0 = time.perf_counter()
    nets.ve_embed_wavs(sdv, wavs)
    nets.campplus_embed_wavs(sdc, wavs)
    dt = time.perf_counter() - t0
    return len(wavs) / dt, dt


def parity_sample(sdv, sdc, wavs, pick, ve, xv):
    """The timed configuration's own output against the CPU oracle on a sample of its clips (north_star gate: cos >= 0.9999,
    max-abs <= 1e-3; the x-vector is not normalised, so its absolute error is reported beside the error relative to max|x|)."""
    from oracle import nets
    want_ve = nets.ve_embed_wavs(sdv, [wavs[i] for i in pick])
    want_xv = nets.campplus_embed_wavs(sdc, [wavs[i] for i in pick])
    cos = lambda a, b: float(np.dot(a.astype(np.float64), b.astype(np.float64)) / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b.astype(np.float64))))
    gv, gx = ve[pick], xv[pick]
    return {"clips_checked": [int(i) for i in pick], "max_abs_ve": float(np.abs(gv - want_ve).max()),
            "max_abs_xv": float(np.abs(gx - want_xv).max()), "max_abs_xv_over_max_x": float(np.abs(gx - want_xv).max() / max(1e-30, np.abs(want_xv).max())),
            "max_x": float(np.abs(want_xv).max()),
            "min_cos": min(min(cos(a, b) for a, b in zip(gv, want_ve)), min(cos(a, b) for a, b in zip(gx, want_xv))),
            "all_finite": bool(np.isfinite(ve).all() and np.isfinite(xv).all()),
            "ve_unit_norm_max_dev": float(np.abs(np.linalg.norm(ve, axis=1) - 1).max()),
            "oracle": "oracle/nets.py (CPU fp32 restatement pinned to the verbatim reference modules)"}


def tf32_peak(dev, seconds: float = 2.0):
    """Dense TF32 peak measured on this box: torch.matmul (cuBLAS, allow_tf32) 8192^3, best of 10 (burst) and back to back
    for `seconds` (sustained) -- the same protocol MEASURED_PEAKS.json uses for bf16."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev); c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"burst": fl / best / 1e9, "sustained": fl / sus / 1e9, "seconds": e0.elapsed_time(e1) / 1e3,
                "how": f"torch.matmul fp32 with allow_tf32, 8192^3, best of 10 / {reps} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def kernel_profile(ctx, step_fn, prof_steps, dev, rank, world, sync_all):
    """Per-kernel device times of `prof_steps` separate profiled steps -> (roofline of the top kernel family, per-kernel table,
    the TF32 dense peak measured in this run)."""
    import torch
    # ---- per-kernel device times (CUDA events on the launching stream), separate profiled steps ------------------
    # (the two encoder chains are serialised for this pass so that a kernel's events time that kernel alone)
    ctx.set_option("overlap", 0)
    step_fn()                                  # one unrecorded step in the serialised configuration (warm-up of this launch order)
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    for _ in range(prof_steps):
        step_fn()
    torch.cuda.synchronize()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    ctx.set_option("overlap", 1)
    # kernel families: the per-conv tags of the FCM head ("fcm_conv_gemm:l1b0c1" ...) are one kernel
    fam = {}
    for k, v in prof.items():
        f = fam.setdefault(k.split(":")[0], dict(ms=0.0, launches=0, flops=0.0, bytes=0.0, exec_flops=0.0))
        for key in ("ms", "launches", "flops", "bytes"):
            f[key] += v[key]
    tot_ms = sum(v["ms"] for v in fam.values()) or 1.0
    tname, t = max(fam.items(), key=lambda kv: kv[1]["ms"])
    pk = peaks()
    tf32 = tf32_peak(dev) if rank == 0 else None
    if world > 1:
        sync_all()
    # TF32 dense peak: measured on this box in this run (cuBLAS 8192^3); a kernel timed inside the step is held against the
    # sustained figure.  (Round 1 assumed half the measured bf16 rate.)
    tensor_peak = tf32["sustained"] if tf32 else pk["bf16_sus"] / 2.0
    tensor_src = "TF32 dense peak measured in this run (torch.matmul allow_tf32 8192^3, sustained)" if tf32 else "bf16 sustained / 2"
    ridge = tensor_peak * 1e12 / (pk["hbm"] * 1e9)          # FLOP per byte above which a kernel is tensor bound
    traffic = traffic_launch = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram__bytes_read+write per launch from ncu --set full captures
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(tname, {})
        traffic, traffic_launch = rec.get("dram_bytes_per_launch"), rec.get("launch")
    secs = t["ms"] / 1e3
    common = {"kernel": tname, "share_of_step": t["ms"] / tot_ms, "avg_launch_ms": t["ms"] / t["launches"],
              "launches_per_step": t["launches"] / prof_steps, "traffic": traffic,
              # which launch the ncu capture is (its own algorithmic bytes are in the note); `achieved` averages over all launches
              "traffic_launch": traffic_launch, "algorithmic_bytes_per_launch_avg": (t["bytes"] / t["launches"]) if t["launches"] else None,
              "arithmetic_intensity_flop_per_byte": (t["flops"] / t["bytes"]) if t["bytes"] else None}
    if t["bytes"] > 0 and (t["flops"] == 0 or t["flops"] / t["bytes"] < ridge):
        achieved = t["bytes"] / secs / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                "peak_source": f"{pk['src']} HBM copy bandwidth", "tflops": t["flops"] / secs / 1e12, **common}
    elif t["flops"] > 0:
        achieved = t["flops"] / secs / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "peak_source": tensor_src, **common}
    else:
        roof = {"bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, **common}
    # per kernel: ALGORITHMIC flops / bytes per second (what the reference op needs); the 3xTF32 front-ends
    # execute three MMAs per algorithmic one -- `exec_tflops` says what the tensor pipe actually ran
    EXEC_FACTOR = {"kaldi_dftmel_tc_kernel": 3.0, "ve_dftmel_tc_kernel": 3.0}
    kernels = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        tfl = (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] > 0 and v["ms"] > 0 else None
        gbs = (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["bytes"] > 0 and v["ms"] > 0 else None
        kernels[k] = {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps, "tflops": tfl, "gbs": gbs,
                      "frac_hbm": gbs / pk["hbm"] if gbs else None, "frac_tf32": tfl / tensor_peak if tfl else None}
        if k in EXEC_FACTOR and tfl:
            kernels[k]["exec_tflops"] = tfl * EXEC_FACTOR[k]
    return roof, kernels, tf32


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


def setup(args):
    import torch
    import torch.distributed as dist
    from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, scheduler
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                       # random-init weights of the same architecture (no checkpoints offline)
    ve = VoiceEncoder().to(dev).eval()
    cp = CAMPPlus().to(dev).eval()
    emb = scheduler.SpeakerEmbedder(ve, cp)
    ctx = emb.ctx()
    ctx.set_option("mode", args.mode)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    sdv = {k: v.detach().cpu() for k, v in ve.state_dict().items()}
    sdc = {k: v.detach().cpu() for k, v in cp.state_dict().items()}
    return world, rank, local, dev, emb, ctx, sdv, sdc


def precision_fields(args, ctx):
    cat_bf16 = int(ctx.get_option("cat_bf16"))                # opt-in (--opt cat_bf16=1|2): not the parity mode, say so in the line
    xw_bf16 = int(ctx.get_option("xw_bf16"))
    dtype = "f32" if args.mode == 0 else ("tf32" if not (cat_bf16 or xw_bf16) else "tf32+bf16")
    note = None if args.mode == 0 else ("tcgen05 kind::tf32 with fp32 accumulation everywhere; the front-end DFT is 3xTF32 "
                                        "(hi/lo split, fp32-accurate); activations and state are stored in fp32"
                                        + ("" if not cat_bf16 else f"; EXCEPT option cat_bf16={cat_bf16}: the D-TDNN bottleneck / transit GEMMs read a bf16 "
                                           "copy of the concatenation buffers" + (" and run on bf16 operands (kind::f16)" if cat_bf16 == 2 else "")
                                           + " -- a looser-tolerance setting, not the fp32/TF32 parity mode")
                                        + ("" if not xw_bf16 else "; the LSTM input projections are stored as bf16 (bf16 mode, --mode 2: tolerance table in tests/test_gpu_parity.py MODE2_TOL)"))
    return dtype, note


def job_core(args, world, rank, local, dev, emb, ctx, sdv, sdc, n, ragged, warm=1):
    """BASELINE configs[3]: partition -> stream every rank's shard -> one all-gather -> un-permute -> parity sample.  Collective on
    every rank; returns the result dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from chatterbox_embed_b200 import scheduler, synth
    lengths = synth.ragged_lengths(n, seed=77) if ragged else np.full(n, CLIP_SAMPLES, dtype=np.int64)
    max_samples = CLIPS * CLIP_SAMPLES
    # host PCM store: a pinned "tape" of synthetic audio (noise and chirp clips back to back); batch b of rank r is the window
    # of its length starting at a batch-dependent offset, its clips are consecutive slices of that window.  Every batch is copied
    # host->device inside the timed job; the tape stands in for the decoded audio a real job would hold in host memory.
    tape_clips = 3 * CLIPS
    tape = torch.empty(tape_clips * CLIP_SAMPLES, dtype=torch.float32).pin_memory()
    tape_np = tape.numpy()
    for i in range(tape_clips):
        tape_np[i * CLIP_SAMPLES:(i + 1) * CLIP_SAMPLES] = synth.clip(i, CLIP_SAMPLES)
    span = len(tape_np) - max_samples - 30 * 16000
    pin_full = torch.empty((n // world + 1, scheduler.EMB), dtype=torch.float32).pin_memory()     # this rank's embeddings, host side

    def window(r, b, total):
        start = ((r * 1000003 + b * 7919) * 160) % span
        return tape_np[start:start + total]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def job(lengths):
        n = len(lengths)
        t = {}
        t0 = time.perf_counter()
        shards = scheduler.partition(lengths, world)
        t["partition_ms"] = 1e3 * (time.perf_counter() - t0)
        mine = shards[rank]
        my_len = lengths[mine]
        t1 = time.perf_counter()
        pin = pin_full[:len(mine)]
        local_emb, status = scheduler.embed_shard(emb, lambda b, i0, i1: window(rank, b, int(my_len[i0:i1].sum())), my_len,
                                                  max_clips=CLIPS, max_samples=max_samples, pinned=True, out=pin.numpy())
        t["stream_ms"] = 1e3 * (time.perf_counter() - t1)
        t2 = time.perf_counter()
        loc = pin.to(dev, non_blocking=True)
        if world > 1:
            full = scheduler.gather_embeddings(loc, shards, n)
        else:
            full = loc[torch.as_tensor(scheduler.inverse_permutation(shards, n), device=dev)]
        torch.cuda.synchronize()
        t["gather_unpermute_ms"] = 1e3 * (time.perf_counter() - t2)
        t["job_ms"] = 1e3 * (time.perf_counter() - t0)
        return full, shards, status, t

    clk = ClockSampler(local).start()
    # warm-up: a short job (allocations, NCCL communicator, first-launch costs)
    for _ in range(warm):
        job(lengths[:min(n, 4 * CLIPS * world)])
    sync_all()
    l0 = ctx.launch_count()
    tw0 = time.time()
    full, shards, status, t = job(lengths)
    sync_all()
    tw1 = time.time()
    launches = ctx.launch_count() - l0
    clocks = clk.summary(tw0, tw1)
    clk.stop()
    tt = torch.tensor([t["partition_ms"], t["stream_ms"], t["gather_unpermute_ms"], t["job_ms"]], device=dev, dtype=torch.float64)
    bad_t = torch.tensor([int((status != 0).sum())], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(bad_t)
    part_ms, stream_ms, gather_ms, job_ms = [float(x) for x in tt]
    if rank != 0:
        return None
    full_np = full.cpu().numpy()
    # parity: 8 sampled clips (spread over the ranks, shortest and longest included) against the oracle
    rng = np.random.RandomState(5)
    pick = sorted(set([int(np.argmin(lengths)), int(np.argmax(lengths))] + [int(x) for x in rng.choice(n, 6, replace=False)]))
    wavs = {}
    rank_of = np.empty(n, dtype=np.int64)
    for r, sh in enumerate(shards):
        rank_of[sh] = r
    for i in pick:
        r = int(rank_of[i])
        pos = int(np.searchsorted(shards[r], i))
        ln = lengths[shards[r]]
        for b, (i0, i1) in enumerate(scheduler.batch_bounds(ln, CLIPS, max_samples)):
            if i0 <= pos < i1:
                w = window(r, b, int(ln[i0:i1].sum()))
                o = int(ln[i0:pos].sum())
                wavs[i] = np.array(w[o:o + int(lengths[i])])
    order = list(wavs)
    par = parity_sample(sdv, sdc, [wavs[i] for i in order], list(range(len(order))), full_np[order, :256], full_np[order, 256:])
    par["clips_checked"] = order
    par["ranks_of_clips_checked"] = [int(rank_of[i]) for i in order]
    audio_s = float(lengths.sum()) / 16000.0
    return {"value": n / (job_ms / 1e3), "unit": "clips/s", "audio_s_per_s": audio_s / (job_ms / 1e3), "job_clips": n, "audio_seconds": audio_s,
            "kind": "3-30 s ragged" if ragged else "10 s each", "scaling": "strong",
            "partition_ms": part_ms, "stream_ms": stream_ms, "gather_unpermute_ms": gather_ms, "job_ms": job_ms,
            "clips_per_rank": [int(len(sh)) for sh in shards], "clips_with_status": int(bad_t[0]),
            "collective": "one all_gather_into_tensor of the padded (clips/rank, 448) fp32 block (NCCL)" if world > 1 else None,
            "timing": "host wall clock per phase, max over ranks; the job is bracketed by barrier + cuda synchronize",
            "h2d_bytes": int(lengths.sum()) * 4, "d2h_bytes": n * (448 + 1) * 4, "gpu_launches": int(launches), "clocks": clocks, "parity": par}


def run_job(args):
    """--config 4: the sharded voice-bank job as the bench line (strong scaling)."""
    import torch.distributed as dist
    world, rank, local, dev, emb, ctx, sdv, sdc = setup(args)
    j = job_core(args, world, rank, local, dev, emb, ctx, sdv, sdc, args.job_clips, args.ragged, warm=max(1, min(args.warmup, 2)))
    if rank == 0:
        dtype, note = precision_fields(args, ctx)
        line = {"metric": "speaker embeddings/sec (voice-bank job)", "value": j["value"], "unit": "clips/s", "audio_s_per_s": j["audio_s_per_s"],
                "n_gpus": world, "steps": 1, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": j["job_ms"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "precision_note": note,
                "config": {"workload": WORKLOADS[4].format(n=j["job_clips"], kind=j["kind"]), "job_clips": j["job_clips"],
                           "audio_seconds": j["audio_seconds"], "parallelism": f"dp{world}", "batch_clips": CLIPS,
                           "l2": "every batch (<= 164 MB PCM) and its activations exceed the 126 MB L2; no flush needed"},
                "job": {k: j[k] for k in ("partition_ms", "stream_ms", "gather_unpermute_ms", "job_ms", "clips_per_rank", "clips_with_status", "collective", "timing")},
                "e2e": {"value": j["value"], "unit": "clips/s", "h2d_bytes_per_step": j["h2d_bytes"], "d2h_bytes_per_step": j["d2h_bytes"],
                        "api": "scheduler.partition -> scheduler.embed_shard (embed_stream) -> scheduler.gather_embeddings"},
                "gpu_launches": j["gpu_launches"], "clocks": j["clocks"], "parity": j["parity"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_conditionals(args):
    """--config 5 (BASELINE configs[4]): the whole `prepare_conditionals` speaker-conditioning path over a batch of 24 kHz prompt
    clips, bf16 mode by default -- what tts.py:329-387 / s3gen.py:150-207 do one clip at a time: resample 24 -> 16 kHz, the 24 kHz
    prompt mel (`prompt_feat`), the S3 tokenizer's 128-bin log-mel front end on the 16 kHz audio (the tokenizer network itself is a
    third-party model outside this path), VoiceEncoder + CAMPPlus embeddings, and the first projection each generator applies to
    them (T3CondEnc.spkr_enc, flow.spk_embed_affine_layer).  One GPU; `value` with the 24 kHz PCM resident in HBM, `e2e` from
    pinned host PCM with everything a voice profile stores copied back to the host inside the timed region."""
    import torch
    from chatterbox_embed_b200 import SpeakerProjections, _lib, synth
    from chatterbox_embed_b200.mel import NUM_MELS
    from chatterbox_embed_b200.s3tokenizer import N_MELS as S3_MELS
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"config": 5, "unavailable": "--config 5 is a one-GPU line (clips are independent: the N-GPU figure is the weak-scaling line of --config 2)"}))
        return
    world, rank, local, dev, emb, ctx, sdv, sdc = setup(args)
    torch.manual_seed(1)
    proj = SpeakerProjections().to(dev).eval()
    W, K = max(args.warmup, 3), args.steps
    clk = ClockSampler(local).start()
    SR24, SR16 = 24000, 16000
    n, L24 = CLIPS, 10 * SR24                                  # DEC_COND_LEN = 10 s of 24 kHz audio (tts.py:119)
    wavs24 = [synth.mixed(i, L24) for i in range(n)]
    host = torch.empty(n * L24, dtype=torch.float32).pin_memory()
    for i, w in enumerate(wavs24):
        host[i * L24:(i + 1) * L24] = torch.from_numpy(w)
    L16 = _lib.resample_out_len(SR24, SR16, L24)
    T24, T16 = _lib.prompt_mel_frames(L24), _lib.s3_log_mel_frames(L16)
    off24 = np.arange(n + 1, dtype=np.int64) * L24
    off16 = np.arange(n + 1, dtype=np.int64) * L16
    pcm24 = host.to(dev)
    def buffers():
        return dict(w16=torch.empty(n * L16, dtype=torch.float32, device=dev), pmel=torch.empty((n * T24, NUM_MELS), dtype=torch.float32, device=dev),
                    s3mel=torch.empty(n * S3_MELS * T16, dtype=torch.float32, device=dev))
    B0 = buffers()
    w16, pmel, s3mel = B0["w16"], B0["pmel"], B0["s3mel"]
    res = {}

    def step_device(src=None, B=None, out=None):
        st = torch.cuda.current_stream(dev).cuda_stream
        x = pcm24 if src is None else src
        B = B0 if B is None else B
        out = res if out is None else out
        ctx.resample(x.data_ptr(), off24, SR24, SR16, B["w16"].data_ptr(), off16, st)
        ctx.prompt_mel(x.data_ptr(), off24, B["pmel"].data_ptr(), st)
        ctx.s3_log_mel(B["w16"].data_ptr(), off16, B["s3mel"].data_ptr(), st)
        ve_o, xv_o, status = emb.embed_device(B["w16"], off16)
        out.update(ve=ve_o, xv=xv_o, status=status, t3=proj.t3_speaker_cond(ve_o), flow=proj.flow_speaker_cond(xv_o))

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), time.perf_counter() - t0

    for _ in range(W):
        step_device()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    tw0 = time.time()
    ms, _ = timed(step_device, K)
    tw1 = time.time()
    launches = ctx.launch_count() - l0
    clocks = clk.summary(tw0, tw1)
    clk.stop()
    value = n * K / (ms / 1e3)
    got = {k: v.detach().cpu().numpy() for k, v in res.items()}
    got["prompt_feat"] = pmel.view(n, T24, NUM_MELS).cpu().numpy()
    got["s3_mel"] = s3mel.view(n, S3_MELS, T16).cpu().numpy()

    # ---- end to end: pinned host PCM in, everything a voice profile stores back on the host ----------------------------------
    # two steps in flight on two streams, each with its own device and pinned buffers: the copies of one step overlap the kernels of
    # the other (libcbx orders its own work across streams; the shared workspace is never used by two steps at once)
    slots = []
    for _ in range(2):
        # (every destination is a contiguous pinned tensor of its own: a copy into a strided host view goes through an unpinned
        # temporary and blocks the submitting thread until the step has finished -- the two steps then never overlap)
        slots.append(dict(stream=torch.cuda.Stream(dev), dst24=torch.empty_like(pcm24), B=buffers(), out={},
                          ve=torch.empty((n, 256), dtype=torch.float32).pin_memory(), xv=torch.empty((n, 192), dtype=torch.float32).pin_memory(),
                          t3=torch.empty((n, 1, 1024), dtype=torch.float32).pin_memory(), flow=torch.empty((n, 80), dtype=torch.float32).pin_memory(),
                          mel=torch.empty((n * T24, NUM_MELS), dtype=torch.float32).pin_memory(),
                          status=torch.empty(n, dtype=torch.int32).pin_memory(), busy=False))

    def submit(sl):
        with torch.cuda.stream(sl["stream"]):
            sl["dst24"].copy_(host, non_blocking=True)
            step_device(sl["dst24"], sl["B"], sl["out"])
            o = sl["out"]
            for key in ("ve", "xv", "t3", "flow", "status"):
                sl[key].copy_(o[key], non_blocking=True)
            sl["mel"].copy_(sl["B"]["pmel"], non_blocking=True)
        sl["busy"] = True

    def steps_host(k):
        for i in range(k):
            sl = slots[i & 1]
            if sl["busy"]:
                sl["stream"].synchronize()                     # the caller reads this slot's profile tensors before reusing it
            submit(sl)
        for sl in slots:
            sl["stream"].synchronize(); sl["busy"] = False
    steps_host(2)
    _, wall = timed(lambda: steps_host(K), 1)
    e2e_value = n * K / wall
    last = slots[(K - 1) & 1]
    e2e_same = float(max(np.abs(last["ve"].numpy() - got["ve"]).max(), np.abs(last["xv"].numpy() - got["xv"]).max(),
                         np.abs(last["t3"].numpy() - got["t3"]).max(), np.abs(last["flow"].numpy() - got["flow"]).max(),
                         np.abs(last["mel"].numpy().reshape(got["prompt_feat"].shape) - got["prompt_feat"]).max()))
    d2h_bytes = sum(int(last[k].numel() * last[k].element_size()) for k in ("ve", "xv", "t3", "flow", "mel", "status"))

    roof, kernels, tf32 = kernel_profile(ctx, step_device, min(K, 2), dev, 0, 1, torch.cuda.synchronize)

    # ---- parity of the timed batch against the CPU oracle, and the CPU port timed on the host cores -----------------------------
    from oracle import frontend, nets
    lin = lambda x, l: torch.nn.functional.linear(torch.from_numpy(x), l.weight.detach().cpu(), l.bias.detach().cpu()).numpy()

    def oracle_clip(w24):
        w16o = frontend.resample_torchaudio(w24, SR24, SR16)
        ve_o = nets.ve_embed_wavs(sdv, [w16o])[0]
        xv_o = nets.campplus_embed_wavs(sdc, [w16o])[0]
        return dict(w16=w16o, prompt_feat=frontend.prompt_mel_torch(w24), s3_mel=frontend.s3_log_mel_torch(w16o), ve=ve_o, xv=xv_o,
                    t3=lin(ve_o[None], proj.spkr_enc)[0],
                    flow=lin((xv_o / max(np.linalg.norm(xv_o), 1e-12))[None], proj.spk_embed_affine_layer)[0])
    pick = [0, 1, n // 2, n - 1]
    cosf = lambda a, b: float(np.dot(a.astype(np.float64).ravel(), b.astype(np.float64).ravel()) / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b.astype(np.float64))))
    par = {"clips_checked": pick, "max_abs": {}, "min_cos": {}}
    w16_np = w16.view(n, L16).cpu().numpy()
    for i in pick:
        o = oracle_clip(wavs24[i])
        o_pm = np.asarray(o["prompt_feat"]).reshape(T24, NUM_MELS)
        pairs = {"resampled_16k": (w16_np[i], np.asarray(o["w16"]).reshape(-1)), "prompt_feat": (got["prompt_feat"][i], o_pm),
                 "s3_log_mel": (got["s3_mel"][i], np.asarray(o["s3_mel"]).reshape(S3_MELS, -1)), "ve": (got["ve"][i], o["ve"]), "xv": (got["xv"][i], o["xv"]),
                 "t3_spkr_enc": (got["t3"][i].reshape(-1), o["t3"]), "flow_affine": (got["flow"][i], o["flow"])}
        for kname, (a, b) in pairs.items():
            par["max_abs"][kname] = max(par["max_abs"].get(kname, 0.0), float(np.abs(a - b).max()))
            par["min_cos"][kname] = min(par["min_cos"].get(kname, 1.0), cosf(a, b))
    par["all_finite"] = bool(all(np.isfinite(v).all() for v in got.values()))
    par["status_nonzero"] = int((got["status"] != 0).sum())
    par["e2e_equals_device_max_abs"] = e2e_same
    par["oracle"] = "oracle/frontend.py + oracle/nets.py (CPU fp32 restatement pinned to the verbatim reference modules / torchaudio)"
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        n_s = 64
        oracle_clip(wavs24[0])
        t0 = time.perf_counter()
        for i in range(n_s):
            oracle_clip(wavs24[i])
        dt = time.perf_counter() - t0
        cpu = {"value": n_s / dt, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": f"{n_s} of the {n} ten-second 24 kHz prompts, B=1 loop, {dt:.1f} s of CPU work (oracle port: torchaudio-style resampler, prompt mel, S3 log-mel, both encoders, projections)"}
    dtype, note = precision_fields(args, ctx)
    line = {"metric": "prepare_conditionals speaker conditionings/sec (10 s 24 kHz prompts)", "value": value, "unit": "clips/s",
            "audio_s_per_s": value * 10.0, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "precision_note": note,
            "config": {"workload": WORKLOADS[5], "mode": {0: "strict-fp32 SIMT", 1: "tcgen05 TF32", 2: "bf16 mode (tcgen05 TF32 + bf16 D-TDNN GEMM operands + bf16 LSTM input projections)"}[args.mode],
                       "l2": f"inputs ({n * L24 * 4 / 1e6:.0f} MB of 24 kHz PCM per step) and activations exceed the 126 MB L2; no flush needed",
                       "parallelism": "dp1", "clips_per_gpu": n, "audio_seconds_per_gpu": n * 10.0},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": n * L24 * 4,
                    "d2h_bytes_per_step": d2h_bytes,
                    "api": "cbx_resample + cbx_prompt_mel + cbx_s3_log_mel + cbx_embed + cbx_project on torch's stream; pinned host PCM in, "
                           "embeddings, projections and prompt_feat back to pinned host memory; two steps in flight on two streams"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": par,
            "tf32_peak_tflops": tf32, "timed_region_s": ms / 1e3, "kernels": kernels}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", type=int, default=None, help="0 strict fp32, 1 tcgen05 TF32 (default), 2 bf16 mode (default of --config 5)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs, 1-based: 2 = 256 x 10 s (default), 3 = ragged 1024, 4 = voice-bank job, 5 = prepare_conditionals (bf16 mode unless --mode is given)")
    ap.add_argument("--job-clips", type=int, default=100000)
    ap.add_argument("--ragged", action="store_true", help="--config 4: 3-30 s clips instead of 10 s")
    ap.add_argument("--opt", action="append", default=[], help="libcbx option key=value (cbx_set_option)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain", type=float, default=3.0, help="seconds of back-to-back steps for the `sustained` object of the line (0 = skip)")
    ap.add_argument("--no-job", action="store_true", help="N > 1: skip the 1e5-clip voice-bank job that is attached to the line")
    args = ap.parse_args()
    if args.mode is None:
        args.mode = int(os.environ.get("CBX_MODE", "2" if args.config == 5 else "1"))
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        return run_job(args)
    if args.config == 5:
        return run_conditionals(args)

    import torch
    import torch.distributed as dist
    from chatterbox_embed_b200 import _lib, scheduler

    world, rank, local, dev, emb, ctx, sdv, sdc = setup(args)
    W = max(args.warmup, 3)
    K = args.steps
    clk = ClockSampler(local).start()          # early: the first nvidia-smi line takes a few hundred ms

    wavs, off = make_batch(rank, args.config)
    n_clips = len(wavs)
    total = int(off[-1])
    host = torch.empty(total, dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    for i, w in enumerate(wavs):
        host_np[off[i]:off[i + 1]] = w
    pcm = host.to(dev)
    gathered = torch.empty((world * n_clips, scheduler.EMB), dtype=torch.float32, device=dev) if world > 1 else None
    dev_out = {}

    def step_device():
        ve_o, xv_o, status = emb.embed_device(pcm, off)
        dev_out["ve"], dev_out["xv"], dev_out["status"] = ve_o, xv_o, status
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
    value = world * n_clips * K / (ms / 1e3)
    clk.stop()
    ve_dev = dev_out["ve"].cpu().numpy(); xv_dev = dev_out["xv"].cpu().numpy(); st_dev = dev_out["status"].cpu().numpy()

    # ---- end to end through host buffers (cbx_embed_host) -------------------------------------------------------
    flags_pinned = _lib.DO_VE | _lib.DO_XV | _lib.PCM_PINNED
    out_holder = {}
    host_block = torch.empty((n_clips, scheduler.EMB), dtype=torch.float32).pin_memory()

    def publish(ve_o, xv_o):
        """N>1: the whole job includes returning every rank's embeddings to every rank (one all-gather per step)."""
        out_holder["ve"], out_holder["xv"] = ve_o, xv_o
        if world > 1:
            hb = host_block.numpy()
            hb[:, :256] = ve_o; hb[:, 256:] = xv_o
            dist.all_gather_into_tensor(gathered, host_block.to(dev))

    def step_host():
        ve_o, xv_o, status = ctx.embed_host(host_np, off, 20.0, 77, 0.8, flags_pinned)
        publish(ve_o, xv_o)

    for _ in range(2):
        step_host()
    _, wall_sync = timed(step_host, K)
    # the voice-bank call a user makes for many batches: SpeakerEmbedder.embed_stream, two batches in flight.  Every step
    # still copies its own PCM host->device and its embeddings device->host inside the timed region; the copies
    # of batch k+1 overlap the kernels of batch k.
    def steps_stream(k):
        n_done = 0
        for ve_o, xv_o, status in emb.embed_stream(((host_np, off) for _ in range(k)), pinned=True):
            publish(ve_o, xv_o)
            n_done += 1
        assert n_done == k
    steps_stream(2)
    _, wall = timed(lambda: steps_stream(K), 1)
    e2e_value = world * n_clips * K / wall
    e2e_sync_value = world * n_clips * K / wall_sync

    roof, kernels, tf32 = kernel_profile(ctx, step_device, min(K, 2), dev, rank, world, sync_all)

    # ---- the same step repeated for >= --sustain seconds (AFTER every other measurement, so that it does not pre-heat them): the
    # figure a long job sees once the clocks have settled under the power cap
    sustained = None
    if args.sustain > 0 and args.config == 2:
        clk2 = ClockSampler(local).start()
        k_s = max(K, int(args.sustain * 1e3 / (ms / K)) + 1)
        for _ in range(2):
            step_device()
        ts0 = time.time()
        ms_s, _ = timed(step_device, k_s)
        ts1 = time.time()
        sustained = {"value": world * n_clips * k_s / (ms_s / 1e3), "unit": "clips/s", "steps": k_s, "seconds": ms_s / 1e3,
                     "ms_per_step": ms_s / k_s, "clocks": clk2.summary(ts0, ts1)}
        clk2.stop()

    # ---- N > 1: BASELINE configs[3] as well -- the 1e5-clip voice-bank job (strong scaling) on the same ranks ------------------
    job = None
    if world > 1 and args.config == 2 and not args.no_job:
        job = job_core(args, world, rank, local, dev, emb, ctx, sdv, sdc, args.job_clips, False, warm=1)
    if rank == 0:
        # ---- parity of the timed configuration itself -------------------------------------------------------------
        lens = np.diff(off)
        pick = sorted({0, 1, n_clips // 2, n_clips - 1, int(np.argmin(lens)), int(np.argmax(lens))})
        parity = parity_sample(sdv, sdc, wavs, pick, ve_dev, xv_dev)
        parity["status_nonzero"] = int((st_dev != 0).sum())
        parity["e2e_equals_device_max_abs"] = float(max(np.abs(out_holder["ve"] - ve_dev).max(), np.abs(out_holder["xv"] - xv_dev).max()))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            if args.config == 3:
                sample = wavs[:64]
                v, dt = cpu_baseline(len(sample), threads, sample)
                cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                       "sample": f"the first 64 of the 1024 ragged clips ({sum(len(w) for w in sample) / 16000:.0f} s of audio), B=1 loop, {dt:.1f} s of CPU work (oracle port, torch CPU fp32)"}
            else:
                n_s = 192                          # ~15-20 s of CPU work on the box's host cores
                v, dt = cpu_baseline(n_s, threads)
                cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                       "sample": f"{n_s} of the {CLIPS} ten-second clips, B=1 loop, {dt:.1f} s of CPU work (oracle port, torch CPU fp32)"}
        dtype, note = precision_fields(args, ctx)
        audio_s = total / 16000.0
        line = {
            "metric": "speaker embeddings/sec (10 s clips)" if args.config == 2 else "speaker embeddings/sec (ragged 3-30 s clips)",
            "value": value, "unit": "clips/s",
            "audio_s_per_s": world * audio_s * K / (ms / 1e3), "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "precision_note": note,
            "config": {"workload": WORKLOADS[args.config], "mode": {0: "strict-fp32 SIMT", 1: "tcgen05 TF32", 2: "bf16 mode (tcgen05 TF32 + bf16 D-TDNN GEMM operands + bf16 LSTM input projections)"}[args.mode],
                       "l2": f"inputs ({total * 4 / 1e6:.0f} MB PCM per step) and activations exceed the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world}", "clips_per_gpu": n_clips, "audio_seconds_per_gpu": audio_s},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": total * 4,
                    "d2h_bytes_per_step": n_clips * (256 + 192 + 1) * 4,
                    "api": "SpeakerEmbedder.embed_stream (cbx_embed_host_submit/_wait, two batches in flight)"
                           + ("; one all_gather_into_tensor of the embeddings per step inside the timed region" if world > 1 else ""),
                    "single_call_value": e2e_sync_value, "single_call_api": "cbx_embed_host (copy, compute, copy back, sync)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity, "job": job, "sustained": sustained,
            "tf32_peak_tflops": tf32, "timed_region_s": ms / 1e3,
            "algorithmic_tflops": (value * FLOPS_PER_CLIP / 1e12) if args.config == 2 else None, "kernels": kernels,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
