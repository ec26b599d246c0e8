"""Training-step timing: WaveGlow.forward + WaveGlowLoss + backward + Adam on BASELINE.json
configs[3]'s shape (batch 32 x 16000-sample segments, 63 mel frames, config.json architecture, weight-norm layout).

    python tools/bench_train.py [--batch 32] [--samples 16000] [--steps 5] [--warmup 2]

Prints one JSON line: ms per step (CUDA events), split into forward / backward / optimiser, audio samples per second,
and the tensor-core GEMM rate (algorithmic FLOPs: forward 522.19 MFLOP per group step, backward 2x that).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import text2speech_b200 as t2s                                   # noqa: E402
from text2speech_b200 import synthetic as syn                    # noqa: E402
from text2speech_b200.training import FusedAdam, allreduce_gradients   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--samples", type=int, default=16000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--graph", action="store_true", help="replay the whole step from one CUDA graph (GraphedTrainStep)")
    ap.add_argument("--breakdown", action="store_true", help="one extra step with CUDA events around every C-ABI call")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = syn.load_config()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = t2s.WaveGlow(**cfg)
    model.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01, weight_norm=True))
    model = model.to(dev).train()
    opt = FusedAdam(model.parameters(), lr=1e-4)
    crit = t2s.WaveGlowLoss(1.0)
    frames = args.samples // 256 + 1
    g = torch.Generator().manual_seed(1)
    audio = (0.1 * torch.randn((args.batch, args.samples), generator=g)).clamp(-1, 1).to(dev)
    mel = syn.synthetic_mel(args.batch, frames, seed=0).to(dev)
    if args.graph:
        from text2speech_b200.training import GraphedTrainStep
        step = GraphedTrainStep(model, opt, crit, args.batch, mel.shape[1], frames, args.samples)
        losses = []
        for _ in range(args.warmup):
            step(mel, audio)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            losses.append(step(mel, audio).clone())
        e1.record()
        torch.cuda.synchronize()
        total = e0.elapsed_time(e1) / args.steps
        t = args.samples // 8
        flop_fwd = 522_190_848 * t * args.batch
        print(json.dumps({
            "metric": "waveglow_train_step_ms", "value": total, "unit": "ms", "higher_is_better": False,
            "config": {"workload": f"WaveGlow train step, batch {args.batch} x {args.samples} samples ({frames} frames), "
                                   "config.json arch, weight norm, Adam; whole step replayed from one CUDA graph"},
            "samples_per_s": args.batch * args.samples / (total * 1e-3),
            "wn_gemm_tflops": {"step": 3 * flop_fwd / (total * 1e-3) / 1e12},
            "loss_first_last": [float(losses[0]), float(losses[-1])], "steps": args.steps, "warmup": args.warmup,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
        }))
        return
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    losses = []
    for it in range(args.warmup + args.steps):
        e = ev[it - args.warmup] if it >= args.warmup else None
        opt.zero_grad()
        if e: e[0].record()
        loss = crit(model((mel, audio)))
        if e: e[1].record()
        loss.backward()
        if e: e[2].record()
        scale = allreduce_gradients(opt)
        opt.step(grad_scale=scale, gathered=True)
        if e: e[3].record()
        losses.append(loss.detach())          # no host sync inside the loop: the next step's launches overlap this one's kernels
    torch.cuda.synchronize()
    losses = [float(v) for v in losses]
    fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    optim = sum(e[2].elapsed_time(e[3]) for e in ev) / args.steps
    total = fwd + bwd + optim
    breakdown = None
    if args.breakdown:
        from text2speech_b200 import _lib
        raw_call, events = _lib.call, []

        def timed_call(name, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw_call(name, *a)
            e1.record()
            events.append((name, e0, e1))

        _lib.call = timed_call
        opt.zero_grad()
        loss = crit(model((mel, audio)))
        loss.backward()
        opt.step(grad_scale=allreduce_gradients(opt), gathered=True)
        torch.cuda.synchronize()
        _lib.call = raw_call
        agg = {}
        for name, a, b in events:
            n, ms = agg.get(name, (0, 0.0))
            agg[name] = (n + 1, ms + a.elapsed_time(b))
        breakdown = {k: {"launches": n, "total_ms": round(ms, 3), "avg_ms": round(ms / n, 4)}
                     for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])}
    t = args.samples // 8
    flop_fwd = 522_190_848 * t * args.batch
    print(json.dumps({
        "metric": "waveglow_train_step_ms", "value": total, "unit": "ms", "higher_is_better": False,
        "config": {"workload": f"WaveGlow train step, batch {args.batch} x {args.samples} samples ({frames} frames), "
                               "config.json arch, weight norm, Adam"},
        "forward_ms": fwd, "backward_ms": bwd, "optimizer_ms": optim,
        "samples_per_s": args.batch * args.samples / (total * 1e-3),
        "wn_gemm_tflops": {"forward": flop_fwd / (fwd * 1e-3) / 1e12, "backward": 2 * flop_fwd / (bwd * 1e-3) / 1e12,
                           "step": 3 * flop_fwd / (total * 1e-3) / 1e12},
        "loss_first_last": [losses[0], losses[-1]], "steps": args.steps, "warmup": args.warmup,
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
        "breakdown": breakdown,
    }))


if __name__ == "__main__":
    main()
