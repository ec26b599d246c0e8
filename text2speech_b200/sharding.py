"""Utterance sharding for the multi-GPU vocoding path (one process per GPU, no collective on the
data path: utterances are independent — SURVEY §8e).  torch.distributed is plumbing only: a
barrier / MAX-reduction for timing and an optional gather of the finished audio to rank 0.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_items`` utterances owned by ``rank``; ragged counts give the
    first ``n_items % world`` ranks one extra item; ranks beyond the item count get an empty slice."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """MAX of a per-rank scalar (device timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_utterances(local: torch.Tensor, n_items: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Collect per-rank outputs [n_local, ...] into [n_items, ...] on rank ``dst`` in utterance order
    (None elsewhere).  Shards may be ragged or empty; this is outside the timed hot path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    widest = max(shard_bounds(n_items, r, world)[1] - shard_bounds(n_items, r, world)[0] for r in range(world))
    pad = torch.zeros((widest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bucket = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bucket, dst=dst)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_items, r, world)
        parts.append(bucket[r][: hi - lo])
    return torch.cat(parts, dim=0)


class MultiGpuVocoder:
    """Single-process form of the utterance sharding: one WaveGlow replica per visible GPU, one host thread per replica.

    ``bench.py`` / torchrun use one PROCESS per GPU; this class is the convenience for callers that want one Python
    process (a server, a notebook): ``infer(mel [B,80,F], sigma, z)`` splits the batch with ``shard_bounds``, runs the
    shards concurrently (the C ABI is safe to call from different host threads on different devices) and concatenates
    the audio on the host.  No inter-GPU traffic: utterances are independent.
    """

    def __init__(self, model, devices=None):
        import copy
        if devices is None:
            devices = [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
        if not devices:
            raise RuntimeError("MultiGpuVocoder needs at least one CUDA device; there is no CPU fallback")
        self.devices = [torch.device(d) for d in devices]
        self.replicas = []
        for i, d in enumerate(self.devices):
            # WaveGlow.__getstate__ makes the copy carry parameters only (no packed-weight cache of device 0, no
            # weight_norm non-leaf tensors), so weight-normed checkpoints replicate too
            replica = model if i == 0 else copy.deepcopy(model)
            self.replicas.append(replica.to(d).eval())

    def infer(self, spect: torch.Tensor, sigma: float = 1.0, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """spect [B, n_mel, F] (host or any device) -> audio [B, 256 F] on the host."""
        import threading
        n = spect.shape[0]
        world = len(self.devices)
        out = [None] * world
        errors = []

        def work(r):
            try:
                lo, hi = shard_bounds(n, r, world)
                if hi == lo:
                    return
                d = self.devices[r]
                with torch.cuda.device(d), torch.no_grad():
                    zz = None if z is None else z[lo:hi].to(d, non_blocking=True)
                    audio = self.replicas[r].infer(spect[lo:hi].to(d, non_blocking=True), sigma=sigma, z=zz)
                    out[r] = audio.cpu()
            except Exception as e:      # noqa: BLE001  (re-raised on the caller's thread)
                errors.append(e)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return torch.cat([o for o in out if o is not None], dim=0)
