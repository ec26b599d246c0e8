"""Multi-process host logic on CPU (gloo, world_size 2): utterance sharding, ragged gather, MAX-reduced
timings, and the reference arm of bench.py running on rank 0 only."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from text2speech_b200.sharding import gather_utterances, max_over_ranks, shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 5, 8, 64, 67):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(64, 3, 8) == (24, 32)
    with pytest.raises(ValueError):
        shard_bounds(4, 4, 4)


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_items, rank, world)
    # stand-in for the per-rank audio: row u of utterance u is filled with u
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 6)
    full = gather_utterances(local, n_items, dst=0)
    slowest = max_over_ranks(10.0 + rank)
    if rank == 0:
        ok = full is not None and full.shape == (n_items, 6) and torch.equal(full[:, 0], torch.arange(n_items).float())
        with open(os.path.join(out_dir, "r0.json"), "w") as f:
            json.dump({"ok": bool(ok), "slowest": slowest}, f)
    else:
        assert full is None and slowest == 10.0 + world - 1
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 1, 8])
def test_gather_and_max_world2(tmp_path, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    res = json.load(open(tmp_path / "r0.json"))
    assert res["ok"] and res["slowest"] == 11.0


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_multi_gpu_vocoder_refuses_cpu_only_hosts():
    from text2speech_b200.sharding import MultiGpuVocoder
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        MultiGpuVocoder(torch.nn.Identity())


def _grad_worker(rank, world, port, out_dir):
    """Data-parallel gradient reduction of the training direction (waveglow/distributed.py:90-142): every rank ends up
    with the SUM in its flat gradient buffer and the 1 / world_size scale to hand to the optimiser."""
    from text2speech_b200.training import allreduce_gradients
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FlatOptimizer:                       # the two members allreduce_gradients uses of FusedAdam
        def __init__(self):
            self.grad = torch.zeros(10)

        def gather_grads(self):
            self.grad.copy_(torch.arange(10, dtype=torch.float32) * (rank + 1))
            return self.grad

    opt = FlatOptimizer()
    scale = allreduce_gradients(opt)
    want = torch.arange(10, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(opt.grad, want) and abs(scale - 1.0 / world) < 1e-12
    with open(os.path.join(out_dir, f"g{rank}.json"), "w") as f:
        json.dump({"ok": bool(ok)}, f)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(json.load(open(tmp_path / f"g{r}.json"))["ok"] for r in range(2))


def _reducer_worker(rank, world, port, out_dir):
    """_GradReducer (per-flow buckets of effective-weight gradients, reduced while the backward is still running) and
    apply_gradient_allreduce's broadcast, on gloo: every rank ends with the MEAN of the ranks' gradients, in views of
    the right shapes, for overlap on and off; parameters start identical to rank 0's."""
    from text2speech_b200 import training
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for overlap in (True, False):
        red = training._GradReducer(None, overlap)
        grads, want = {}, {}
        for k in range(3):                                   # three "flows" with a few tensors each
            names = []
            for j, shape in enumerate([(4, 3, 1), (7,), (2, 5)]):
                n = f"WN.{k}.t{j}"
                base = torch.arange(float(torch.Size(shape).numel())).reshape(shape) + 10 * k + j
                grads[n] = base * (rank + 1)
                want[n] = base * sum(r + 1 for r in range(world)) / world
                names.append(n)
            red.add(grads, names)
        red.finish(grads)
        ok = ok and all(grads[n].shape == want[n].shape and torch.allclose(grads[n], want[n]) for n in want)
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 1))
    training.apply_gradient_allreduce(lin)
    ok = ok and bool((lin.weight == 1.0).all()) and lin._dp_allreduce == (None, True)
    with open(os.path.join(out_dir, f"r{rank}.json"), "w") as f:
        json.dump({"ok": bool(ok)}, f)
    dist.destroy_process_group()


def test_per_flow_gradient_reducer_world2(tmp_path):
    port = _free_port()
    mp.spawn(_reducer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(json.load(open(tmp_path / f"r{r}.json"))["ok"] for r in range(2))
