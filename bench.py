#!/usr/bin/env python
"""Benchmark of the vocoding hot path: WaveGlow.infer audio samples/s on B200(s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the samples/s metric is quoted on): 64 utterances of
80x860 mel frames (10 s at 22.05 kHz, 220 160 samples each), sigma 0.666, config.json architecture,
random-init weights, host-supplied noise; the batch is sharded by utterance over the N ranks with
no collective on the data path (total work fixed -> "strong" scaling).

One JSON line on rank 0.  `value` = whole-job samples/s with mel and z already resident in HBM;
`e2e` = the same through the public module API from pinned host buffers (H2D of mel+z, D2H of the
audio inside the timed region); `roofline` = the gate GEMM (tcgen05 in_layers+cond kernel), its
algorithmic FLOPs / CUDA-event time measured inside the timed region, against the measured bf16
peak; `cpu_baseline` = the CPU oracle port of the reference timed on this box's host cores on a
bounded sample (one 80x100 mel).  `--impl reference` prints the CPU arm on its own (the reference is
pure Python + PyTorch CPU ops; /root/reference cannot travel to the GPU box, so the oracle port —
the same torch CPU ops in the same order — is what is timed).
"""
from __future__ import annotations

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

METRIC = "waveglow_infer_audio_samples_per_sec"
UNIT = "samples/s"
GLOBAL_BATCH = 64
FRAMES = 860
SIGMA = 0.666
SAMPLE_RATE = 22050
CPU_SAMPLE_FRAMES = 100
PUBLISHED_V100_SAMPLES_PER_SEC = 2.75e6      # BASELINE.md §1 (waveglow/README.md:15-16, 1x V100 fp16)
GATE_FLOP_PER_STEP = 2 * (3 * 512 + 640) * 1024   # in_layers + cond_layers MACs*2 per group step per layer
GATE_ENTRY_POINTS = ("wgb_tc_wn_gate", "wgb_tc2_wn_gate", "wgb_tc2_wn_gate_mel")
# dram__bytes_read.sum + dram__bytes_write.sum of one gate-GEMM launch at the full per-GPU batch of 64
# (ncu --set full, profiles/r01c_ncu_full_summary.csv: 6.667 + 1.791 GB); algorithmic bytes are
# h 1 KB + cond 1.25 KB read + acts 1 KB written per group step = 5.84 GB.  Scales with the per-rank batch.
GATE_DRAM_BYTES_PER_LAUNCH_B64 = {"wgb_tc2_wn_gate": 8.458e9,         # profiles/r01c_ncu_full_summary.csv
                                  "wgb_tc2_wn_gate_mel": 4.16e9}      # profiles/r01t_ncu_full_summary.csv (2.38 + 1.78 GB
                                                                      # for the launch captured there; 2.24 + 1.78 GB in r01l)
GATE_DRAM_SOURCE = "profiles/r01t_ncu_full_summary.csv (composed gate) / r01c (cond-tensor gate)"


def workload_config(n_gpus):
    return {
        "workload": "WaveGlow.infer 64 x (80x860 mel, 10 s @22.05 kHz), sigma 0.666, config.json arch, random init "
                    "(BASELINE.json configs[2])",
        "global_batch": GLOBAL_BATCH, "frames": FRAMES, "samples_per_utt": FRAMES * 256,
        "parallelism": f"utterance-sharded x{n_gpus}, no collective on the data path",
        "l2": "no flush needed: per-step working set (~20 GB activations) >> 126 MB L2",
    }


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["bf16_tflops"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


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

def cpu_reference_time(steps, warmup):
    """Oracle port of the reference's CPU path on one 80x100 mel; returns (samples/s, cores, sample text)."""
    import torch
    import oracle
    from text2speech_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.synthetic_state_dict(syn.load_config(), seed=1234, end_std=0.01)
    mel = syn.synthetic_mel(1, CPU_SAMPLE_FRAMES, seed=0)
    z = syn.synthetic_z(1, CPU_SAMPLE_FRAMES, seed=2024)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.waveglow_infer(sd, mel, z, SIGMA)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    sample = (f"oracle port of WaveGlow.infer, 1 x 80x{CPU_SAMPLE_FRAMES} mel ({CPU_SAMPLE_FRAMES * 256} samples), fp32, "
              f"torch {torch.__version__} CPU, {torch.get_num_threads()} threads, mean of {len(times)} runs "
              f"({sec:.2f} s each); cost is linear in batch x frames")
    return CPU_SAMPLE_FRAMES * 256 / sec, cores, sample, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    value, cores, sample, sec = cpu_reference_time(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = pure-Python/PyTorch CPU path; /root/reference is not present on the GPU box, so the "
                "oracle port (same torch CPU ops, pinned to the reference by tests/golden) is timed",
    }
    print(json.dumps(line), flush=True)


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

    # ---- launch counter + per-launch CUDA events on the dominant kernel (gate GEMM)
    counter = {"n": 0}
    gate_events = []
    gate_names = []
    seen_entry_points = set()
    raw_call = _lib.call
    profile = {"on": False}

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
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw_call(name, *a)
            e1.record()
            gate_events.append((e0, e1))
            gate_names.append(name)
        else:
            raw_call(name, *a)

    _lib.call = counted_call

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
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        return model.infer(mel_dev, sigma=SIGMA, z=z_dev)

    def step_e2e():
        m = mel_host.to(dev, non_blocking=True)
        zz = z_host.to(dev, non_blocking=True)
        out_host.copy_(model.infer(m, sigma=SIGMA, z=zz), non_blocking=True)

    for _ in range(args.warmup):
        step_resident()
    counter["n"] = 0
    profile["on"] = True
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(step_resident, args.steps)
    profile["on"] = False
    launches = counter["n"]
    gate_ms = [a.elapsed_time(b) for a, b in gate_events]
    step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    breakdown = None
    if args.breakdown:                       # one extra step with CUDA events around every C-ABI call
        profile["all"] = True
        step_resident()
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
    wn_flop_total = 522190848 * GLOBAL_BATCH * t_steps       # WN GEMM FLOPs per group step
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
    traffic = GATE_DRAM_BYTES_PER_LAUNCH_B64.get(gate_name)
    roofline = {"kernel": kernel_desc, "entry_point": gate_name,
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": (achieved / sustained) if achieved else None, "frac_of_burst": (achieved / burst) if achieved else None,
                "peak_source": peak_src + ", bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": len(gate_ms), "avg_launch_ms": gate_avg_ms,
                "flop_per_launch": gate_flop, "executed_flop_per_launch": gate_exec,
                "executed_tflops": achieved_exec, "frac_executed": (achieved_exec / sustained) if achieved_exec else None,
                "traffic": traffic * per_rank / GLOBAL_BATCH if traffic else None,
                "traffic_source": "ncu dram__bytes_read+write per launch at batch 64, " + GATE_DRAM_SOURCE,
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

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_reference_time(2, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / PUBLISHED_V100_SAMPLES_PER_SEC, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "rtf": (ms_per_step * 1e-3) / (samples_total / SAMPLE_RATE),
            "wn_gemm_tflops_per_gpu": overall_tflops,
            "wn_gemm_frac_of_bf16_peak": {"sustained": overall_tflops / sustained, "burst": overall_tflops / burst},
            "wn_gemm_executed": executed,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": (mel_host.numel() + z_host.numel()) * 4 * world,
                    "d2h_bytes_per_step": out_host.numel() * 4 * world},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "breakdown": breakdown,
            "clocks": clocks.summary(),
            "baseline_note": "vs_baseline divides by the 2750 kHz (1x V100, fp16) figure of waveglow/README.md:15-16",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="add per-entry-point CUDA-event totals of one extra step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
