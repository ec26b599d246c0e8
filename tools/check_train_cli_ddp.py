"""Data-parallel run of the training CLI (text2speech_b200.train = waveglow/train.py + distributed.py) on synthetic wavs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
        tools/check_train_cli_ddp.py /tmp/t2s_ddp

Rank 0 writes eight short wav files, the file list and a reference-format config.json (4 flows to keep it quick); every
rank then runs ``train.main(["-c", config])`` (env:// rendezvous, DistributedSampler, flat gradient all-reduce) for two
epochs and rank 0 checks the checkpoint it wrote.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text2speech_b200 import synthetic as syn, train as t2s_train        # noqa: E402


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/tmp/t2s_ddp"
    rank = int(os.environ.get("RANK", 0))
    cfg_path = os.path.join(root, "config.json")
    if rank == 0:
        from scipy.io.wavfile import write
        os.makedirs(root, exist_ok=True)
        g = torch.Generator().manual_seed(3)
        names = []
        for i in range(8):
            path = os.path.join(root, f"u{i}.wav")
            write(path, 22050, (0.2 * torch.randn(9000 + 300 * i, generator=g)).clamp(-1, 1).mul(32767).short().numpy())
            names.append(path)
        with open(os.path.join(root, "files.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        wg = dict(syn.load_config())
        wg.update(n_flows=4, n_early_every=2, n_early_size=2)
        config = {
            "train_config": {"output_directory": os.path.join(root, "ckpt"), "epochs": 2, "learning_rate": 1e-5, "sigma": 1.0,
                             "iters_per_checkpoint": 1, "batch_size": 2, "seed": 1234, "checkpoint_path": ""},
            "data_config": {"training_files": os.path.join(root, "files.txt"), "segment_length": 4096, "sampling_rate": 22050,
                            "filter_length": 1024, "hop_length": 256, "win_length": 1024, "mel_fmin": 0.0, "mel_fmax": 8000.0},
            "dist_config": {"dist_backend": "nccl", "dist_url": "tcp://localhost:54321"},
            "waveglow_config": wg,
        }
        with open(cfg_path + ".tmp", "w") as f:
            json.dump(config, f)
        os.replace(cfg_path + ".tmp", cfg_path)
    else:
        for _ in range(600):
            if os.path.exists(cfg_path):
                break
            time.sleep(0.1)
    t2s_train.main(["-c", cfg_path])
    if rank == 0:
        ck = torch.load(os.path.join(root, "ckpt", "waveglow_3"), map_location="cpu", weights_only=False)
        ok = ck["iteration"] == 3 and int(ck["optimizer"]["state"][0]["step"]) == 4
        print(json.dumps({"check": "train_cli_ddp", "world": int(os.environ.get("WORLD_SIZE", 1)), "ok": bool(ok),
                          "iterations": ck["iteration"] + 1}), flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
