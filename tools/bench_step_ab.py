"""Step-level A/B of a wgb_set_tuning switch: the headline step (WaveGlow.infer, 64 x 80x860 mel) captured twice as a CUDA
graph -- once per value of the switch, which is a kernel parameter baked in at capture time -- and replayed alternately in
ONE process (box-to-box clock variance is larger than the effects looked for), with NVML power / SM clock per block.

    python tools/bench_step_ab.py [--key gate_l2_hint] [--values 0,1] [--batch 64] [--rounds 4] [--steps 6]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import text2speech_b200 as t2s                           # noqa: E402
from text2speech_b200 import _lib, synthetic as syn      # noqa: E402
from tools.bench_kernels import Nvml                     # noqa: E402

DEV = torch.device("cuda:0")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--key", default="gate_l2_hint")
    ap.add_argument("--values", default="0,1")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=860)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_ab.json"))
    args = ap.parse_args()
    values = [int(v) for v in args.values.split(",")]
    cfg = syn.load_config()
    model = t2s.WaveGlow.remove_weightnorm(t2s.WaveGlow(**cfg))
    model.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01))
    model = model.to(DEV).eval()
    mel = syn.synthetic_mel(args.batch, args.frames, seed=0).to(DEV)
    z = syn.synthetic_z(args.batch, args.frames, seed=2024).to(DEV)
    runs, outs = {}, {}
    for v in values:
        _lib.call("wgb_set_tuning", args.key, v)
        runs[v] = model.graphed_infer(args.batch, args.frames, sigma=0.666)
        outs[v] = runs[v](mel, z).clone()
    torch.cuda.synchronize()
    same = all(torch.equal(outs[values[0]], outs[v]) for v in values)
    recs = []
    for r in range(args.rounds):
        for v in values:
            runs[v].replay()                              # one untimed replay after the switch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with Nvml() as nv:
                e0.record()
                for _ in range(args.steps):
                    runs[v].replay()
                e1.record()
                e1.synchronize()
            rec = {"round": r, args.key: v, "ms_per_step": round(e0.elapsed_time(e1) / args.steps, 3), **nv.summary()}
            recs.append(rec)
            print(json.dumps(rec), flush=True)
    summary = {}
    for v in values:
        ms = sorted(x["ms_per_step"] for x in recs if x[args.key] == v)
        summary[str(v)] = {"median_ms": ms[len(ms) // 2], "min_ms": ms[0], "max_ms": ms[-1]}
    out = {"key": args.key, "batch": args.batch, "frames": args.frames, "steps_per_block": args.steps,
           "outputs_bit_identical": same, "summary": summary, "blocks": recs}
    print(json.dumps({"summary": summary, "outputs_bit_identical": same}))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
