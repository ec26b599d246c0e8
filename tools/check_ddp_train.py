"""Data-parallel training check on N GPUs of one box (waveglow/train.py + waveglow/distributed.py:90-142):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/check_ddp_train.py [--batch 8] [--samples 16000] [--steps 3]

Every rank trains the same model on its own shard and Adam runs on every rank.  --reduce picks how the gradients meet:
  flat      ONE NCCL all-reduce of the optimiser's flat gradient buffer after the backward pass (allreduce_gradients)
  overlap   apply_gradient_allreduce(model): one all-reduce per flow of the effective-weight gradients, issued as soon as
            that flow's backward is done, i.e. hidden behind the remaining flows (the reference's name and contract)
  deferred  the same buckets, all issued after the last flow (what `overlap` is compared against)
Checks that the parameters stay bit-identical across ranks and prints step times (device-timed, max over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import warnings

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import text2speech_b200 as t2s                                   # noqa: E402
from text2speech_b200 import synthetic as syn                    # noqa: E402
from text2speech_b200.training import FusedAdam, allreduce_gradients, apply_gradient_allreduce   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch")
    ap.add_argument("--samples", type=int, default=16000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="forward + backward + gradient gather from one CUDA graph per rank")
    ap.add_argument("--reduce", default="flat", choices=["flat", "overlap", "deferred"])
    args = ap.parse_args()
    if args.graph and args.reduce != "flat":
        raise SystemExit("--graph goes with --reduce flat (per-flow collectives captured into the graph were measured slower "
                         "and hang the process group at teardown: profiles/r02g_*)")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = syn.load_config()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = t2s.WaveGlow(**cfg)
    model.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01, weight_norm=True))
    model = model.to(dev).train()
    if args.reduce != "flat":
        model = apply_gradient_allreduce(model, overlap=args.reduce == "overlap")
    opt = FusedAdam(model.parameters(), lr=1e-6)
    crit = t2s.WaveGlowLoss(1.0)
    frames = args.samples // 256 + 1
    g = torch.Generator().manual_seed(100 + rank)                 # every rank its own shard of the global batch
    audio = (0.1 * torch.randn((args.batch, args.samples), generator=g)).clamp(-1, 1).to(dev)
    mel = syn.synthetic_mel(args.batch, frames, seed=200 + rank).to(dev)
    step_ms, ar_ms, losses = [], [], []
    graphed = None
    if args.graph:
        from text2speech_b200.training import GraphedTrainStep
        graphed = GraphedTrainStep(model, opt, crit, args.batch, mel.shape[1], frames, args.samples, include_optimizer=False)
    for it in range(args.steps + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        dist.barrier()
        torch.cuda.synchronize()
        e[0].record()
        if graphed is not None:
            loss = graphed(mel, audio)                            # ends with the flat gradient gathered
            e[1].record()
            scale = allreduce_gradients(opt, gathered=True) if args.reduce == "flat" else None
        else:
            opt.zero_grad()
            loss = crit(model((mel, audio)))
            loss.backward()                                       # overlap / deferred: gradients leave averaged
            e[1].record()
            scale = allreduce_gradients(opt) if args.reduce == "flat" else (opt.gather_grads(), 1.0)[1]
        e[2].record()
        if scale is not None:
            opt.step(grad_scale=scale, gathered=True)
        e[3].record()
        torch.cuda.synchronize()
        if it > 0:                                                # first iteration = warm-up (NCCL setup, allocator)
            step_ms.append(e[0].elapsed_time(e[3]))
            ar_ms.append(e[1].elapsed_time(e[2]))
        losses.append(float(loss.detach()))
    # parameters must be identical on every rank after identical updates
    flat = opt.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(flat, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    t = torch.tensor([sum(step_ms) / len(step_ms), sum(ar_ms) / len(ar_ms)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "ddp_train", "world": world, "per_gpu_batch": args.batch, "samples": args.samples,
                          "graph": bool(args.graph), "reduce": args.reduce, "params_identical_across_ranks": bool(same.item() == 1.0),
                          "step_ms": float(t[0]), "grad_gather_plus_allreduce_ms": float(t[1]),
                          "flat_gradient_bytes": opt.n * 4,
                          "samples_per_s_all_gpus": world * args.batch * args.samples / (float(t[0]) * 1e-3),
                          "rank0_losses": losses}), flush=True)
    ok = same.item() == 1.0
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
