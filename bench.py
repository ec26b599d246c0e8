#!/usr/bin/env python
"""Benchmark of the vocoding hot path: WaveGlow.infer audio samples/s on B200(s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the samples/s metric is quoted on): 64 utterances of
80x860 mel frames (10 s at 22.05 kHz, 220 160 samples each), sigma 0.666, config.json architecture,
random-init weights, host-supplied noise; the batch is sharded by utterance over the N ranks with
no collective on the data path (total work fixed -> "strong" scaling).

One JSON line on rank 0.
  value         whole-job samples/s with mel and z already resident in HBM.  A step is one replay of the CUDA graph
                that `WaveGlow.graphed_infer` (public API) captures of `infer` for this shape (--no-graph: eager calls).
  e2e           the same through the public API from pinned host buffers: H2D of mel + z, the replay, D2H of the audio,
                all inside the timed region.
  roofline      the gate GEMM (tcgen05 in_layers + composed conditioning kernel): algorithmic FLOPs / CUDA-event time
                of its launches inside the timed region (event-record nodes captured with the graph), against the
                measured bf16 peak; `traffic` is read from the newest profiles/*_ncu_full_summary.csv.
  cpu_baseline  the CPU oracle port of the reference timed on this box's host cores on a bounded sample (N = 1 only).
  gpu_eager_baseline   the same oracle port (plain PyTorch ops = the reference's op sequence) run on THIS GPU through
                cuDNN / cuBLAS in fp32 (TF32 convs) and bf16: "the practical kernel to beat" of SURVEY §2.1 / §6.
  secondary     the other BASELINE.json configs: cfg2 single-utterance latency / RTF (eager and graph), cfg4 forward
                32 x 16 000, cfg5 mel + denoiser on 256 x 10 s waveforms (sharded over the ranks like the headline).
`--impl reference` prints the CPU arm on its own (the reference is pure Python + PyTorch CPU ops; /root/reference
cannot travel to the GPU box, so the oracle port -- the same torch CPU ops in the same order, pinned to the
reference's outputs by tests/golden -- is what is timed).
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "waveglow_infer_audio_samples_per_sec"
UNIT = "samples/s"
GLOBAL_BATCH = 64
FRAMES = 860
SIGMA = 0.666
SAMPLE_RATE = 22050
CPU_SAMPLE_FRAMES = 100                      # BASELINE.json configs[0]: the reference's own CPU-runnable case
CPU_BUDGET_S = 150.0
PUBLISHED_V100_SAMPLES_PER_SEC = 2.75e6      # BASELINE.md §1 (waveglow/README.md:15-16, 1x V100 fp16)
WN_FLOP_PER_STEP = 522190848                 # in + cond + res_skip GEMM FLOPs per group step, all 12 flows
GATE_FLOP_PER_STEP = 2 * (3 * 512 + 640) * 1024   # in_layers + cond_layers MACs*2 per group step per layer
GATE_ENTRY_POINTS = ("wgb_tc_wn_gate", "wgb_tc2_wn_gate", "wgb_tc2_wn_gate_mel")
GATE_KERNEL_OF_ENTRY = {"wgb_tc2_wn_gate_mel": "pair_kernel<3, 0, 0>", "wgb_tc2_wn_gate": "pair_kernel<0, 0, 0>"}
STFT_FLOP_PER_FRAME = 2 * 1026 * 1024        # dense-basis conv, one direction
MEL_FLOP_PER_FRAME = 2 * 80 * 513


def workload_config(n_gpus):
    return {
        "workload": "WaveGlow.infer 64 x (80x860 mel, 10 s @22.05 kHz), sigma 0.666, config.json arch, random init "
                    "(BASELINE.json configs[2])",
        "global_batch": GLOBAL_BATCH, "frames": FRAMES, "samples_per_utt": FRAMES * 256,
        "parallelism": f"utterance-sharded x{n_gpus}, no collective on the data path",
        "l2": "no flush needed: per-step working set (~20 GB activations + 2.6 GB weights) >> 126 MB L2",
    }


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["bf16_tflops"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_name, per_rank_batch):
    """dram__bytes_read + dram__bytes_write of one launch of `kernel_name` from the NEWEST
    profiles/*_ncu_full_summary.csv that holds it (tools/ncu_summary.py output of `ncu --set full` on this bench at the
    full per-GPU batch of 64), scaled to this rank's batch.  Several captured launches of one kernel (the first-layer
    variant shares the template instance): the longest launch is the K = 1856 layer.  Returns (bytes, source) or
    (None, reason)."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_summary.csv")), reverse=True):
        try:
            with open(path) as f:
                rows = list(csv.reader(f))
        except OSError:
            continue
        if len(rows) < 3 or "dram__bytes_read.sum" not in rows[0]:
            continue
        hdr, units = rows[0], rows[1]
        i_name, i_t = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for r in rows[2:]:
            if kernel_name in r[i_name]:
                t = float(r[i_t])
                b = float(r[i_r]) * scale.get(units[i_r], 1.0) + float(r[i_w]) * scale.get(units[i_w], 1.0)
                if best is None or t > best[0]:
                    best = (t, b, os.path.relpath(path, ROOT))
        if best is not None:
            break
    if best is None:
        return None, "no profiles/*_ncu_full_summary.csv holds " + kernel_name
    return best[1] * per_rank_batch / GLOBAL_BATCH, (f"{best[2]}: dram__bytes_read.sum + dram__bytes_write.sum of the "
                                                     f"longest captured launch of {kernel_name} (batch 64), scaled to "
                                                     f"this rank's batch of {per_rank_batch}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm

def cpu_reference_time(steps, warmup, budget_s=CPU_BUDGET_S):
    """Oracle port of the reference's CPU path, one 80x100 mel (BASELINE.json configs[0]) per step; stops early when
    the time budget is spent.  Returns (samples/s, cores, sample text, s per step, timed steps, warm-ups)."""
    import torch
    import oracle
    from text2speech_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.synthetic_state_dict(syn.load_config(), seed=1234, end_std=0.01)
    mel = syn.synthetic_mel(1, CPU_SAMPLE_FRAMES, seed=0)
    z = syn.synthetic_z(1, CPU_SAMPLE_FRAMES, seed=2024)
    times, warm_done = [], 0
    t_start = time.perf_counter()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.waveglow_infer(sd, mel, z, SIGMA)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            else:
                warm_done += 1
            if times and time.perf_counter() - t_start > budget_s:
                break
    sec = sum(times) / len(times)
    sample = (f"oracle port of WaveGlow.infer, 1 x 80x{CPU_SAMPLE_FRAMES} mel ({CPU_SAMPLE_FRAMES * 256} samples, "
              f"BASELINE.json configs[0]) per step, fp32, torch {torch.__version__} CPU, {torch.get_num_threads()} threads, "
              f"mean of {len(times)} steps after {warm_done} warm-ups ({sec:.2f} s each); samples/s = samples of the "
              f"sample / time (cost is linear in batch x frames; the full 64 x 80x860 batch would take ~8 min per step)")
    return CPU_SAMPLE_FRAMES * 256 / sec, cores, sample, sec, len(times), warm_done


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, cores, sample, sec, steps, warmup = cpu_reference_time(max(1, args.steps), max(0, args.warmup))
    cfg = workload_config(args.gpus)
    cfg["workload"] += (f" -- reference arm: CPU, bounded sample of that workload = 1 x 80x{CPU_SAMPLE_FRAMES} mel "
                        f"({CPU_SAMPLE_FRAMES * 256} samples) per step")
    cfg["sample_per_step"] = {"batch": 1, "frames": CPU_SAMPLE_FRAMES, "samples": CPU_SAMPLE_FRAMES * 256}
    cfg["parallelism"] = f"host CPU only, {cores} torch threads on rank 0 (the GPUs are idle)"
    cfg["l2"] = "n/a (CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup, "time_budget_s": CPU_BUDGET_S,
        "note": "reference = pure-Python/PyTorch CPU path; /root/reference is not present on the GPU box, so the "
                "oracle port (same torch CPU ops, pinned to the reference by tests/golden) is timed; ms_per_step is "
                "the time of ONE 80x100 sample, not of the 64 x 80x860 batch",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm: baselines / secondaries

def gpu_eager_baseline(dev, hbm_free_gb):
    """The reference's op sequence in plain PyTorch ON THIS GPU (cuDNN / cuBLAS sm_100 kernels): the oracle port run on
    cuda tensors -- conv1d / conv_transpose1d / tanh / sigmoid / einsum exactly as waveglow/glow.py:251-292 issues
    them -- in fp32 (cuDNN TF32 convs, torch's default) and in bf16 (the analogue of waveglow/inference.py:40-43's
    .half() mode; W^-1 computed in fp64 as in the oracle).  Outside every timed region of the product arm; the oracle
    is used as a BASELINE here, never as part of the product path."""
    import torch
    import oracle
    from text2speech_b200 import synthetic as syn
    out = {"kind": "port-on-cuda", "what": "oracle port (the reference's PyTorch op sequence) on cuda:  cuDNN/cuBLAS eager",
           "frames": FRAMES, "unit": UNIT}
    batch = 8 if hbm_free_gb > 60 else 2
    sd = {k: v.to(dev) for k, v in syn.synthetic_state_dict(syn.load_config(), seed=1234, end_std=0.01).items()}
    mel = syn.synthetic_mel(batch, FRAMES, seed=0).to(dev)
    z = syn.synthetic_z(batch, FRAMES, seed=2024).to(dev)
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True            # let cuDNN pick its best algorithm per conv shape
    runs = {}
    try:
        for name, dt in (("fp32_tf32", torch.float32), ("bf16", torch.bfloat16)):
            try:
                sdt = {k: v.to(dt) for k, v in sd.items()}
                m, zz = mel.to(dt), z.to(dt)
                with torch.no_grad():
                    for _ in range(2):
                        oracle.waveglow_infer(sdt, m, zz, SIGMA)
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    iters = 3
                    e0.record()
                    for _ in range(iters):
                        audio = oracle.waveglow_infer(sdt, m, zz, SIGMA)
                    e1.record()
                    torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / iters
                runs[name] = {"value": batch * FRAMES * 256 / (ms * 1e-3), "ms_per_call": ms, "batch": batch,
                              "finite": bool(torch.isfinite(audio.float()).all()),
                              "wn_gemm_tflops": WN_FLOP_PER_STEP * batch * FRAMES * 32 / (ms * 1e-3) / 1e12}
                del sdt, m, zz, audio
            except Exception as e:  # noqa: BLE001   (a baseline must not take the product line down)
                runs[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = prev
    out["runs"] = runs
    ok = [r for r in runs.values() if "value" in r]
    if ok:
        best = max(ok, key=lambda r: r["value"])
        out.update({"value": best["value"], "batch": best["batch"],
                    "dtype": [k for k, r in runs.items() if r is best][0],
                    "note": "value = the faster of the two modes; cudnn.benchmark on, cudnn TF32 on (torch default), "
                            "2 warm-up + 3 timed calls, CUDA events; batch 8 of the 64 utterances (eager keeps "
                            "[B,1024,T] fp32 intermediates: 0.9 GB each at batch 8)"})
    return out


def time_calls(torch, fn, dev, reduce_max, warmup=3, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return reduce_max(e0.elapsed_time(e1) / iters)


def secondary_configs(model, dev, rank, world, barrier, reduce_max, peak):
    """BASELINE.json configs other than the headline, timed after it (CUDA events, 3 warm-ups, mean of 5 back-to-back
    calls, max over ranks).  Between iterations nothing useful survives in L2: every WaveGlow call streams 2.6 GB of
    weights, cfg5's input alone is 225 MB."""
    import torch
    import text2speech_b200 as t2s
    from text2speech_b200 import synthetic as syn
    from text2speech_b200.sharding import shard_bounds
    sustained, burst, hbm = peak
    out = {"timing": "CUDA events, 3 warm-ups, mean of 5 back-to-back calls, max over ranks; L2: every call streams "
                     "more than the 126 MB L2 (2.6 GB of weights / 225 MB of waveforms)"}
    dc = syn.DEFAULT_DATA_CONFIG

    # ---- cfg2: one 10 s utterance, latency and real-time factor (every rank runs the same replica; max reported)
    mel = syn.synthetic_mel(1, FRAMES, seed=0).to(dev)
    z = syn.synthetic_z(1, FRAMES, seed=2024).to(dev)
    samples = FRAMES * 256
    barrier()
    ms = time_calls(torch, lambda: model.infer(mel, sigma=SIGMA, z=z), dev, reduce_max)
    rec = {"config": "BASELINE.json configs[1]: WaveGlow.infer 1 x 80x860 mel (10 s), bf16, 1 GPU (replicas at N > 1)",
           "latency_ms": ms, "rtf": (ms * 1e-3) / (samples / SAMPLE_RATE), "samples_per_s": samples / (ms * 1e-3),
           "wn_gemm_tflops": WN_FLOP_PER_STEP * FRAMES * 32 / (ms * 1e-3) / 1e12}
    rec["frac_bf16_burst"] = rec["wn_gemm_tflops"] / burst
    try:
        run = model.graphed_infer(1, FRAMES, sigma=SIGMA)
        ref = model.infer(mel, sigma=SIGMA, z=z)
        same = bool(torch.equal(run(mel, z), ref))
        gms = time_calls(torch, run.replay, dev, reduce_max)
        rec.update({"graph_latency_ms": gms, "graph_rtf": (gms * 1e-3) / (samples / SAMPLE_RATE),
                    "graph_wn_gemm_tflops": WN_FLOP_PER_STEP * FRAMES * 32 / (gms * 1e-3) / 1e12,
                    "graph_frac_bf16_burst": WN_FLOP_PER_STEP * FRAMES * 32 / (gms * 1e-3) / 1e12 / burst,
                    "graph_matches_eager": same})
        del run
    except Exception as e:  # noqa: BLE001
        rec["graph_error"] = repr(e)[:300]
    out["cfg2_single_utterance"] = rec

    # ---- cfg4: forward direction (audio -> z, log_s, log_det_W), 32 x 16 000 samples sharded by segment
    lo, hi = shard_bounds(32, rank, world)
    g = torch.Generator().manual_seed(1)
    wav = (0.1 * torch.randn((32, 16000), generator=g)).clamp(-1, 1)[lo:hi].contiguous().to(dev)
    taco = t2s.TacotronSTFT(dc["filter_length"], dc["hop_length"], dc["win_length"], 80, dc["sampling_rate"],
                            dc["mel_fmin"], dc["mel_fmax"]).to(dev)
    barrier()
    if hi > lo:
        mel4 = taco.mel_spectrogram(wav)
        ms = time_calls(torch, lambda: model((mel4, wav)), dev, reduce_max)
    else:
        ms = reduce_max(0.0)
    tf = WN_FLOP_PER_STEP * 32 * 2000 / (ms * 1e-3) / 1e12
    out["cfg4_forward"] = {"config": "BASELINE.json configs[3]: WaveGlow.forward 32 x 16 000 samples (mel 63 frames), "
                                     f"bf16, segment-sharded x{world}", "ms": ms, "samples_per_s": 32 * 16000 / (ms * 1e-3),
                           "wn_gemm_tflops": tf, "frac_bf16_sustained_per_gpu": tf / world / sustained}

    # ---- cfg5: mel_spectrogram + Denoiser(0.01) on 256 x 10 s waveforms, sharded by waveform
    lo, hi = shard_bounds(256, rank, world)
    y = syn.synthetic_waveforms(256, FRAMES * 256, sr=dc["sampling_rate"], seed=5)[lo:hi].contiguous().to(dev)
    n_total = 256 * FRAMES * 256
    frames_total = 256 * (FRAMES + 1)
    den = t2s.Denoiser(model)
    mel_flop = frames_total * (STFT_FLOP_PER_FRAME + MEL_FLOP_PER_FRAME)
    den_flop = frames_total * 2 * STFT_FLOP_PER_FRAME
    runs = {}
    for algo in ("auto", "gemm"):      # default = butterfly kernels (csrc/fft.cu); 'gemm' = tensor-core dense-basis kernels
        taco.stft_fn.algorithm = den.stft.algorithm = algo
        barrier()
        ms_mel = time_calls(torch, lambda: taco.mel_spectrogram(y), dev, reduce_max) if hi > lo else reduce_max(0.0)
        barrier()
        ms_mel_k = time_calls(torch, lambda: taco._mel_spectrogram(y), dev, reduce_max) if hi > lo else reduce_max(0.0)
        barrier()
        ms_den = time_calls(torch, lambda: den(y, strength=0.01), dev, reduce_max) if hi > lo else reduce_max(0.0)
        runs[algo] = (ms_mel, ms_mel_k, ms_den)
    taco.stft_fn.algorithm = den.stft.algorithm = "auto"
    fft_used = taco.stft_fn._fft_pack(dev) is not None
    ms_mel, ms_mel_k, ms_den = runs["auto"]
    g_mel, g_mel_k, g_den = runs["gemm"]
    out["cfg5_mel"] = {
        "config": f"BASELINE.json configs[4]: TacotronSTFT.mel_spectrogram, 256 x 220 160 samples, waveform-sharded x{world}",
        "kernel": "wgb_fft_stft_mel: 1024-point real FFT per warp + |X| + mel filterbank + log-clamp, one launch"
                  if fft_used else "wgb_tc2_stft_mel",
        "bound": "hbm (shared-memory / issue limited in practice)" if fft_used else "tensor",
        "ms": ms_mel, "ms_without_range_asserts": ms_mel_k, "samples_per_s": n_total / (ms_mel * 1e-3),
        "algorithmic_gb_s": n_total * 5.25 / (ms_mel_k * 1e-3) / 1e9,
        "frac_hbm_per_gpu": n_total * 5.25 / (ms_mel_k * 1e-3) / 1e9 / world / hbm,
        "dense_basis_equivalent_tflops": mel_flop / (ms_mel_k * 1e-3) / 1e12,
        "dense_basis_tensor_core_path": {
            "kernel": "wgb_tc2_stft_mel (STFT.algorithm = 'gemm'): split-bf16 dense-basis GEMM on CTA pairs",
            "ms": g_mel, "ms_without_range_asserts": g_mel_k,
            "algorithmic_tflops": mel_flop / (g_mel_k * 1e-3) / 1e12,
            "frac_bf16_burst_per_gpu": mel_flop / (g_mel_k * 1e-3) / 1e12 / world / burst},
        "note": "ms = the public call (includes the read of the range flag of layers.py:72-73 = one device->host sync); the "
                "fractions use the kernels alone; 5.25 B/sample = fp32 in + 80/256 fp32 mel out"}
    out["cfg5_denoiser"] = {
        "config": f"BASELINE.json configs[4]: Denoiser(strength 0.01), 256 x 220 160 samples, waveform-sharded x{world}",
        "kernel": "wgb_fft_denoise: FFT, spectral subtraction, inverse FFT, overlap-add in registers, one launch"
                  if fft_used else "wgb_tc2_stft_denoise + wgb_tc2_istft_ola",
        "bound": "hbm (shared-memory / issue limited in practice)" if fft_used else "tensor",
        "ms": ms_den, "samples_per_s": n_total / (ms_den * 1e-3),
        "algorithmic_gb_s": n_total * 8 / (ms_den * 1e-3) / 1e9,
        "frac_hbm_per_gpu": n_total * 8 / (ms_den * 1e-3) / 1e9 / world / hbm,
        "dense_basis_equivalent_tflops": den_flop / (ms_den * 1e-3) / 1e12,
        "dense_basis_tensor_core_path": {
            "kernel": "wgb_tc2_stft_denoise + wgb_tc2_istft_ola (STFT.algorithm = 'gemm')",
            "ms": g_den, "algorithmic_tflops": den_flop / (g_den * 1e-3) / 1e12,
            "frac_bf16_burst_per_gpu": den_flop / (g_den * 1e-3) / 1e12 / world / burst}}
    del y, den, taco
    torch.cuda.empty_cache()
    return out


def measure_train(dev, rank, world, steps, warmup, barrier, reduce_max, peak, batch=32, samples=16000):
    """Training direction (waveglow/train.py:108-124) at BASELINE.json configs[3]'s shape PER GPU: forward,
    WaveGlowLoss, backward, Adam on `batch` x `samples`-sample segments, config.json architecture in weight-norm layout.
    One GPU: the whole step replays from one CUDA graph (GraphedTrainStep).  N > 1 (weak scaling, data parallel): the
    graph ends after the gradient gather and is followed by one flat NCCL all-reduce and Adam (the reduction of
    waveglow/distributed.py:90-142).  `e2e` adds the H2D copy of the batch from pinned host memory and a D2H read of the
    loss every step."""
    import warnings
    import torch
    import text2speech_b200 as t2s
    from text2speech_b200 import synthetic as syn
    from text2speech_b200.training import FusedAdam, GraphedTrainStep, allreduce_gradients
    sustained, burst, hbm = peak
    cfg = syn.load_config()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = t2s.WaveGlow(**cfg)
    model.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01, weight_norm=True))
    model = model.to(dev).train()
    if world > 1:                             # identical start on every rank (distributed.py:99-103)
        import torch.distributed as dist
        for p_ in model.state_dict().values():
            dist.broadcast(p_, 0)
    opt = FusedAdam(model.parameters(), lr=1e-6)     # small steps: the loss falls monotonically on the synthetic batch
    crit = t2s.WaveGlowLoss(1.0)
    frames = samples // 256 + 1
    g = torch.Generator().manual_seed(1 + rank)
    audio_h = (0.1 * torch.randn((batch, samples), generator=g)).clamp(-1, 1).pin_memory()
    mel_h = syn.synthetic_mel(batch, frames, seed=rank).pin_memory()
    audio, mel = audio_h.to(dev), mel_h.to(dev)
    loss_h = torch.zeros(1).pin_memory()
    # forward, loss, backward and gradient gather replay from one CUDA graph; alone on a GPU Adam is part of it, under data
    # parallelism the flat gradient is all-reduced (ONE NCCL call over NVLink) and Adam stepped right after the replay --
    # the fastest of the four variants measured (profiles/r02f_*, r02g_*, r02h_*)
    graphed = GraphedTrainStep(model, opt, crit, batch, mel.shape[1], frames, samples, include_optimizer=world == 1)
    if world == 1:
        def step(m, a):
            return graphed(m, a)
        how = "whole step (forward, loss, backward, gradient gather, Adam) replayed from one CUDA graph"
    else:
        def step(m, a):
            loss = graphed(m, a)
            opt.step(grad_scale=allreduce_gradients(opt, gathered=True), gathered=True)
            return loss
        how = ("forward + loss + backward + gradient gather replayed from one CUDA graph, then ONE flat NCCL all-reduce "
               "(1.07 GB, waveglow/distributed.py:105-129) and the fused Adam kernel")

    def timed(fn):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1) / steps)

    losses = []
    ms = timed(lambda: losses.append(step(mel, audio).clone()))

    def e2e_step():
        m, a = mel_h.to(dev, non_blocking=True), audio_h.to(dev, non_blocking=True)
        loss_h.copy_(step(m, a).reshape(1), non_blocking=True)
    ms_e2e = timed(e2e_step)
    torch.cuda.synchronize(dev)
    flop_fwd = WN_FLOP_PER_STEP * (samples // 8) * batch
    tf = 3 * flop_fwd / (ms * 1e-3) / 1e12                        # per GPU: forward + 2x backward (data + weight gradients)
    out = {"config": f"WaveGlow train step (waveglow/train.py:108-124), {batch} x {samples} samples per GPU (BASELINE.json "
                     f"configs[3] shape, {frames} mel frames), config.json arch, weight norm, FusedAdam, x{world} data parallel",
           "how": how, "ms_per_step": ms, "samples_per_s": world * batch * samples / (ms * 1e-3),
           "e2e_ms_per_step": ms_e2e, "e2e_samples_per_s": world * batch * samples / (ms_e2e * 1e-3),
           "h2d_bytes_per_step": (mel_h.numel() + audio_h.numel()) * 4 * world, "d2h_bytes_per_step": 4 * world,
           "wn_gemm_tflops_per_gpu": tf, "frac_bf16_sustained_per_gpu": tf / sustained,
           "flop_accounting": "algorithmic conv FLOPs of the reference: forward 522.19 MFLOP per group step, backward 2x",
           "loss_first_last": [float(losses[0]), float(losses[-1])],
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    del model, opt
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------ GPU arm

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import text2speech_b200 as t2s
    from text2speech_b200 import _lib, synthetic as syn

    from text2speech_b200.sharding import shard_bounds
    lo, hi = shard_bounds(GLOBAL_BATCH, rank, world)       # contiguous utterance shard of this rank
    per_rank = hi - lo
    model = t2s.WaveGlow(**syn.load_config())
    model = t2s.WaveGlow.remove_weightnorm(model)
    model.load_state_dict(syn.synthetic_state_dict(syn.load_config(), seed=1234, end_std=0.01))
    model = model.to(dev).eval()
    model.mode = "bf16"

    mel_host = syn.synthetic_mel(GLOBAL_BATCH, FRAMES, seed=0)[lo: lo + per_rank].contiguous().pin_memory()
    z_host = syn.synthetic_z(GLOBAL_BATCH, FRAMES, seed=2024)[lo: lo + per_rank].contiguous().pin_memory()
    out_host = torch.empty((per_rank, FRAMES * 256), dtype=torch.float32).pin_memory()
    mel_dev, z_dev = mel_host.to(dev), z_host.to(dev)
    samples_total = GLOBAL_BATCH * FRAMES * 256

    # ---- launch counter + per-launch CUDA events on the dominant kernel (gate GEMM).  Under graph capture the events
    # are "external" ones: they become event-record nodes of the graph and are re-recorded by every replay.
    counter = {"n": 0}
    gate_events = []
    gate_names = []
    seen_entry_points = set()
    raw_call = _lib.call
    profile = {"on": False, "external": False}

    breakdown_events = []

    def counted_call(name, *a):
        counter["n"] += 1
        seen_entry_points.add(name)
        if profile.get("all"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw_call(name, *a)
            e1.record()
            breakdown_events.append((name, e0, e1))
        elif profile["on"] and name in GATE_ENTRY_POINTS:
            kw = {"external": True} if profile["external"] else {}
            e0, e1 = torch.cuda.Event(enable_timing=True, **kw), torch.cuda.Event(enable_timing=True, **kw)
            e0.record()
            raw_call(name, *a)
            e1.record()
            gate_events.append((e0, e1))
            gate_names.append(name)
        else:
            raw_call(name, *a)

    _lib.call = counted_call

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    # ---- the step: a CUDA-graph replay of infer (WaveGlow.graphed_infer, public API) or the eager call
    use_graph = not args.no_graph
    launches_per_step = None
    graph_note = None
    run = None
    if use_graph:
        try:
            model._packed(dev)                      # weights packed outside the capture

            def arm():                              # after graphed_infer's eager warm-up, right before the capture
                counter["n"] = 0
                gate_events.clear()
                gate_names.clear()
                profile["on"], profile["external"] = True, True

            run = model.graphed_infer(per_rank, FRAMES, sigma=SIGMA, before_capture=arm)
            launches_per_step = counter["n"]        # C-ABI calls captured = kernel nodes of one replay
            profile["on"] = False
            run(mel_dev, z_dev)                     # inputs resident in the graph's static buffers from here on
        except Exception as e:  # noqa: BLE001
            graph_note = "graph capture failed, eager steps timed instead: " + repr(e)[:200]
            use_graph, run = False, None
            profile["on"] = profile["external"] = False
            gate_events.clear()
            gate_names.clear()

    if use_graph:
        def step_resident():
            run.replay()

        def step_e2e():
            out_host.copy_(run(mel_host, z_host), non_blocking=True)       # H2D into the static buffers, replay, D2H
    else:
        def step_resident():
            return model.infer(mel_dev, sigma=SIGMA, z=z_dev)

        def step_e2e():
            m = mel_host.to(dev, non_blocking=True)
            zz = z_host.to(dev, non_blocking=True)
            out_host.copy_(model.infer(m, sigma=SIGMA, z=zz), non_blocking=True)

    for _ in range(args.warmup):
        step_resident()
    if not use_graph:
        counter["n"] = 0
        profile["on"] = True
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(step_resident, args.steps)
    profile["on"] = False
    launches = launches_per_step * args.steps if use_graph else counter["n"]
    gate_timing = ("event-record nodes inside the replayed graph, read after the last timed step (the launches of that "
                   "step)" if use_graph else "CUDA events around every gate launch of the timed region")
    try:
        gate_ms = [a.elapsed_time(b) for a, b in gate_events]
        if use_graph and (not gate_ms or min(gate_ms) <= 0.0):
            raise RuntimeError("external events carry no timing")
    except Exception as e:  # noqa: BLE001   (driver without timing on event nodes: two eager steps right after)
        gate_timing = "eager steps right after the timed region (event nodes of the graph gave no timing: %r)" % (e,)
        gate_events.clear()
        gate_names.clear()
        profile["on"], profile["external"] = True, False
        for _ in range(2):
            model.infer(mel_dev, sigma=SIGMA, z=z_dev)
        torch.cuda.synchronize()
        profile["on"] = False
        gate_ms = [a.elapsed_time(b) for a, b in gate_events]
    step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    audio_check = bool(torch.isfinite(out_host).all()) and float(out_host.abs().max()) > 0.0
    breakdown = None
    if args.breakdown:                       # one extra eager step with CUDA events around every C-ABI call
        profile["all"] = True
        model.infer(mel_dev, sigma=SIGMA, z=z_dev)
        torch.cuda.synchronize()
        profile["all"] = False
        agg = {}
        for name, a, b in breakdown_events:
            n, ms = agg.get(name, (0, 0.0))
            agg[name] = (n + 1, ms + a.elapsed_time(b))
        breakdown = {k: {"launches": n, "total_ms": round(ms, 3), "avg_ms": round(ms / n, 4)}
                     for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])}

    ms_per_step = total_ms / args.steps
    value = samples_total / (ms_per_step * 1e-3)
    e2e_value = samples_total / (e2e_ms / args.steps * 1e-3)
    sustained, burst, hbm, peak_src = peaks()
    t_steps = FRAMES * 32
    gate_flop = GATE_FLOP_PER_STEP * per_rank * t_steps
    gate_avg_ms = sum(gate_ms) / max(1, len(gate_ms))
    achieved = gate_flop / (gate_avg_ms * 1e-3) / 1e12 if gate_ms else None
    wn_flop_total = WN_FLOP_PER_STEP * GLOBAL_BATCH * t_steps
    overall_tflops = wn_flop_total / (ms_per_step * 1e-3) / 1e12 / world

    # The gate GEMM runs either on the [B,T,640] cond tensor (K = 2176, 128 group steps per tile) or with the
    # conditioning composed with the upsampler (K = 1856, 128 frames x 1 phase per tile; engine.use_mel_path).
    # `achieved` is ALGORITHMIC FLOPs (the reference's in_layers + cond_layers convs) / time; `executed_*` is what
    # the tensor pipe really ran (fewer K chunks, but whole 128-frame tiles).
    gate_name = gate_names[0] if gate_names else None
    if gate_name == "wgb_tc2_wn_gate_mel":
        gate_exec = 2 * (3 * 512 + 320) * 1024 * (-(-per_rank * (FRAMES + 4) // 128) * 128) * 32   # padded frame axis
        kernel_desc = ("tc2::pair_kernel<GATE_MEL>: in_layers k=3 dilated + (cond_layers o upsample) composed, K = 1856, "
                       "+ gate epilogue; tcgen05 cta_group::2")
    else:
        gate_exec = gate_flop
        kernel_desc = "in_layers k=3 dilated + cond 1x1 (K = 2176) + gate epilogue; tcgen05"
    achieved_exec = gate_exec / (gate_avg_ms * 1e-3) / 1e12 if gate_ms else None
    traffic, traffic_src = ncu_traffic(GATE_KERNEL_OF_ENTRY.get(gate_name, "?"), per_rank)
    roofline = {"kernel": kernel_desc, "entry_point": gate_name,
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": (achieved / sustained) if achieved else None, "frac_of_burst": (achieved / burst) if achieved else None,
                "peak_source": peak_src + ", bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": len(gate_ms), "avg_launch_ms": gate_avg_ms, "timing": gate_timing,
                "flop_per_launch": gate_flop, "executed_flop_per_launch": gate_exec,
                "executed_tflops": achieved_exec, "frac_executed": (achieved_exec / sustained) if achieved_exec else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": 3328 * per_rank * t_steps,      # reference op: h + cond in, acts out
                "min_bytes_per_launch_as_executed": 2048 * per_rank * t_steps + 640 * per_rank * (FRAMES + 4)
                if gate_name == "wgb_tc2_wn_gate_mel" else 3328 * per_rank * t_steps}

    # Tensor-pipe FLOPs actually executed per infer on this rank (three exact algebraic
    # compositions execute fewer FLOPs than the reference's convs, which is what `wn_gemm_tflops_per_gpu` counts)
    n_flows, n_layers = 12, 8
    rows = per_rank * t_steps
    if gate_name == "wgb_tc2_wn_gate_mel":
        tiled_rows = (-(-per_rank * (FRAMES + 4) // 128) * 128) * 32
        fold = 1 if "wgb_tc2_wn_gate_mel0" in seen_entry_points else 0
        exec_flop = n_flows * ((n_layers - fold) * 2 * 1856 * 1024 * tiled_rows + fold * 2 * 384 * 1024 * tiled_rows)
    else:
        exec_flop = n_flows * n_layers * GATE_FLOP_PER_STEP * rows
    exec_flop += n_flows * (n_layers - 1) * 2 * 512 * 512 * rows
    exec_flop += n_flows * (2 * 4096 * 16 * rows if "wgb_tc_wn_skip16_end" in seen_entry_points else 2 * 4096 * 512 * rows)
    exec_tflops = exec_flop / (ms_per_step * 1e-3) / 1e12
    executed = {"tflop_per_step_per_gpu": exec_flop / 1e12, "tflops_per_gpu": exec_tflops,
                "frac_of_bf16_sustained": exec_tflops / sustained,
                "note": "conditioning composed with the upsampler, WN.end composed with the skip sum, WN.start folded "
                        "into in_layers[0] (exact; DESIGN.md section 4)"}

    # per-rank step times (the slowest rank sets `value`): shows whether a scaling loss is one slow GPU or all of them
    rank_ms = None
    if world > 1:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(min(args.steps, 5)):
            step_resident()
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1) / min(args.steps, 5)], device=dev, dtype=torch.float64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        rank_ms = [round(float(t.item()), 3) for t in gathered]

    # free the headline's activations before the secondary configs / the eager baseline
    del run
    torch.cuda.empty_cache()

    secondary = None
    if not args.no_secondary:
        try:
            secondary = secondary_configs(model, dev, rank, world, barrier, reduce_max, (sustained, burst, hbm))
            model.repack()                   # drop the packed inference weights (2.6 GB) before the training step
            torch.cuda.empty_cache()
            secondary["train_step"] = measure_train(dev, rank, world, 5, 3, barrier, reduce_max, (sustained, burst, hbm))
        except Exception as e:  # noqa: BLE001
            if world > 1:
                raise
            secondary = dict(secondary or {}, error=repr(e)[:400])
    eager = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            free_gb = torch.cuda.mem_get_info(dev)[0] / 1e9
            eager = gpu_eager_baseline(dev, free_gb)
            if eager.get("value"):
                eager["speedup_of_value_over_it"] = value / eager["value"]
        except Exception as e:  # noqa: BLE001
            eager = {"kind": "port-on-cuda", "error": repr(e)[:400]}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _, _, _ = cpu_reference_time(10, 2, budget_s=25.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        cfg = workload_config(world)
        cfg["step"] = ("one replay of the CUDA graph WaveGlow.graphed_infer captures for this shape" if use_graph
                       else "one eager WaveGlow.infer call")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / PUBLISHED_V100_SAMPLES_PER_SEC, "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "rtf": (ms_per_step * 1e-3) / (samples_total / SAMPLE_RATE),
            "wn_gemm_tflops_per_gpu": overall_tflops,
            "wn_gemm_frac_of_bf16_peak": {"sustained": overall_tflops / sustained, "burst": overall_tflops / burst},
            "wn_gemm_executed": executed,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": (mel_host.numel() + z_host.numel()) * 4 * world,
                    "d2h_bytes_per_step": out_host.numel() * 4 * world, "output_finite_nonzero": audio_check},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": eager,
            "secondary": secondary,
            "rank_ms_per_step": rank_ms,
            "graph_note": graph_note,
            "breakdown": breakdown,
            "clocks": clocks.summary(),
            "baseline_note": "vs_baseline divides by the 2750 kHz (1x V100, fp16) figure of waveglow/README.md:15-16",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train_arm(args):
    """`--mode train`: the training step as the main line (same JSON contract; metric = training samples/s)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sustained, burst, hbm, peak_src = peaks()
    with ClockSampler(local_rank) as clocks:
        r = measure_train(dev, rank, world, args.steps, args.warmup, barrier, reduce_max, (sustained, burst, hbm))
    if rank == 0:
        line = {"metric": "waveglow_train_audio_samples_per_sec", "value": r["samples_per_s"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": r["config"], "step": r["how"],
                           "l2": "no flush needed: a step streams ~30 GB of activations"},
                "e2e": {"value": r["e2e_samples_per_s"], "unit": UNIT, "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": r["d2h_bytes_per_step"]},
                "roofline": {"bound": "tensor", "achieved": r["wn_gemm_tflops_per_gpu"], "peak": sustained, "unit": "TFLOP/s",
                             "frac": r["frac_bf16_sustained_per_gpu"], "traffic": None,
                             "kernel": "whole step (forward + data-gradient + weight-gradient GEMMs), " + r["flop_accounting"],
                             "peak_source": peak_src + ", bf16_tflops_sustained"},
                "cpu_baseline": None, "detail": r, "clocks": clocks.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch-eager-on-this-GPU baseline")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE.json configs 1, 3, 4")
    ap.add_argument("--no-graph", action="store_true", help="time eager infer calls instead of CUDA-graph replays (ncu runs)")
    ap.add_argument("--breakdown", action="store_true", help="add per-entry-point CUDA-event totals of one extra step")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer (default): the headline vocoding benchmark; train: the training step on 32 x 16 000 samples per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.mode == "train":
        run_train_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
