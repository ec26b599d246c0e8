"""Vocoder training CLI (reference: waveglow/train.py + waveglow/distributed.py's init / reduce helpers).

    python -m text2speech_b200.train -c config.json                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 -m text2speech_b200.train -c config.json     # data parallel

Same ``config.json`` sections (train_config / data_config / dist_config / waveglow_config), the same ``train(...)``
signature, per-iteration log line and checkpoint dict (``{'model': <pickled WaveGlow module>, 'iteration', 'optimizer',
'learning_rate'}``, train.py:52-62) as the reference, so a checkpoint written here loads with
``waveglow/inference.py:37`` (``torch.load(path)['model']``) and with ``text2speech_b200.inference.load_waveglow``.
What runs underneath (training.py): forward / backward through this package's kernels, ``FusedAdam`` (one kernel per
step; its ``state_dict`` is ``torch.optim.Adam``'s format), and ``apply_gradient_allreduce`` (distributed.py:90-142's name
and contract) with one NCCL all-reduce per flow, issued while the backward pass of the remaining flows runs.  The dataset's mel is computed on the GPU, hence ``num_workers=0``.
"""
from __future__ import annotations

import argparse
import json
import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from .glow import WaveGlow, WaveGlowLoss
from .mel2samp import Mel2Samp
from .training import FusedAdam, apply_gradient_allreduce


def reduce_tensor(tensor: torch.Tensor, num_gpus: int) -> torch.Tensor:
    """distributed.py:37-41: mean over ranks."""
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= num_gpus
    return rt


def init_distributed(rank: int, num_gpus: int, group_name: str, dist_backend: str, dist_url: str) -> None:
    """distributed.py:43-54; under torchrun (RANK / WORLD_SIZE / MASTER_* in the environment) the env:// rendezvous is
    used instead of ``dist_url``."""
    assert torch.cuda.is_available(), "Distributed mode requires CUDA."
    print("Initializing Distributed")
    torch.cuda.set_device(rank % torch.cuda.device_count())
    if "MASTER_ADDR" in os.environ and "WORLD_SIZE" in os.environ:
        dist.init_process_group(dist_backend, world_size=num_gpus, rank=rank)
    else:
        dist.init_process_group(dist_backend, init_method=dist_url, world_size=num_gpus, rank=rank)


def load_checkpoint(checkpoint_path, model, optimizer):
    """train.py:41-50."""
    assert os.path.isfile(checkpoint_path)
    from .inference import install_glow_alias
    install_glow_alias()                                  # checkpoints written by the reference pickle ``glow.WaveGlow``
    checkpoint_dict = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    iteration = checkpoint_dict["iteration"]
    model_for_loading = checkpoint_dict["model"]
    model.load_state_dict(model_for_loading.state_dict())
    optimizer.load_state_dict(checkpoint_dict["optimizer"])
    print("Loaded checkpoint '{}' (iteration {})".format(checkpoint_path, iteration))
    return model, optimizer, iteration


def save_checkpoint(model, optimizer, learning_rate, iteration, filepath, waveglow_config):
    """train.py:52-62: the whole module is pickled (a fresh copy holding the current weights)."""
    print("Saving model and optimizer state at iteration {} to {}".format(iteration, filepath))
    model_for_saving = WaveGlow(**waveglow_config)
    model_for_saving.load_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    torch.save({"model": model_for_saving, "iteration": iteration, "optimizer": optimizer.state_dict(),
                "learning_rate": learning_rate}, filepath)


def train(num_gpus, rank, group_name, output_directory, epochs, learning_rate, sigma, iters_per_checkpoint, batch_size,
          seed, checkpoint_path, waveglow_config, data_config, dist_config=None, fp16_run=False, with_tensorboard=False,
          max_iterations=None):
    """train.py:64-140.  ``max_iterations`` (extra) stops early.  ``fp16_run`` / ``with_tensorboard`` do not exist in the
    reference tree this package mirrors (its waveglow/train.py and config.json:2-11 have neither; they belong to later upstream
    revisions): they are accepted for config compatibility and must be False (the GEMMs already run in bf16 with fp32
    accumulation and fp32 master weights, which is what upstream's fp16_run buys with apex)."""
    if fp16_run or with_tensorboard:
        raise ValueError("fp16_run / with_tensorboard are not supported")
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    if num_gpus > 1:
        init_distributed(rank, num_gpus, group_name, **(dist_config or {"dist_backend": "nccl", "dist_url": "env://"}))
    criterion = WaveGlowLoss(sigma)
    model = WaveGlow(**waveglow_config).cuda()
    if num_gpus > 1:
        model = apply_gradient_allreduce(model)            # train.py:75-76
    optimizer = FusedAdam(model.parameters(), lr=learning_rate)
    iteration = 0
    if checkpoint_path != "":
        model, optimizer, iteration = load_checkpoint(checkpoint_path, model, optimizer)
        iteration += 1                                    # next iteration is iteration + 1
    trainset = Mel2Samp(**data_config)
    train_sampler = DistributedSampler(trainset) if num_gpus > 1 else None
    train_loader = DataLoader(trainset, num_workers=0, shuffle=False, sampler=train_sampler, batch_size=batch_size,
                              pin_memory=False, drop_last=True)
    if rank == 0:
        if not os.path.isdir(output_directory):
            os.makedirs(output_directory)
            os.chmod(output_directory, 0o775)
        print("output directory", output_directory)
    model.train()
    epoch_offset = max(0, int(iteration / len(train_loader)))
    losses = []
    for epoch in range(epoch_offset, epochs):
        print("Epoch: {}".format(epoch))
        for i, batch in enumerate(train_loader):
            optimizer.zero_grad()
            mel, audio = batch
            mel, audio = mel.cuda(), audio.cuda()
            outputs = model((mel, audio))
            loss = criterion(outputs)
            reduced_loss = reduce_tensor(loss.data, num_gpus).item() if num_gpus > 1 else loss.item()
            loss.backward()                               # gradients leave averaged over the ranks (distributed.py:105-141),
            optimizer.step()                              # reduced per flow while the backward was still running
            print("{}:\t{:.9f}".format(iteration, reduced_loss))
            losses.append(reduced_loss)
            if iteration % iters_per_checkpoint == 0 and rank == 0:
                save_checkpoint(model, optimizer, learning_rate, iteration,
                                "{}/waveglow_{}".format(output_directory, iteration), waveglow_config)
            iteration += 1
            if max_iterations is not None and len(losses) >= max_iterations:
                return _finish(num_gpus, losses)
    return _finish(num_gpus, losses)


def _finish(num_gpus, losses):
    if num_gpus > 1 and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return losses


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--config", type=str, help="JSON file for configuration")
    parser.add_argument("-r", "--rank", type=int, default=None, help="rank of process for distributed")
    parser.add_argument("-g", "--group_name", type=str, default="", help="name of group for distributed")
    args = parser.parse_args(argv)
    with open(args.config) as f:
        config = json.loads(f.read())
    rank = args.rank if args.rank is not None else int(os.environ.get("RANK", 0))
    num_gpus = int(os.environ.get("WORLD_SIZE", 1)) if "WORLD_SIZE" in os.environ else torch.cuda.device_count()
    if num_gpus > 1 and args.group_name == "" and "WORLD_SIZE" not in os.environ:
        print("WARNING: Multiple GPUs detected but no distributed group set")
        print("Only running 1 GPU.  Use torch.distributed.run for multiple GPUs")
        num_gpus = 1
    if num_gpus == 1 and rank != 0:
        raise Exception("Doing single GPU training on rank > 0")
    train(num_gpus, rank, args.group_name, waveglow_config=config["waveglow_config"], data_config=config["data_config"],
          dist_config=config.get("dist_config"), **config["train_config"])


if __name__ == "__main__":
    main()
