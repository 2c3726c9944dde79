"""CPU tests of the training path: the oracle restatement against the golden step made by the reference itself,
and the data-parallel host logic over gloo with world_size 2 (no GPU needed)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import train_port
from oracle.torch_port import ClassifierPort, load_numpy_state
from tests.util import golden, golden_keep, sample_positions, sha, synth, train_inputs

train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")


def test_oracle_training_step_matches_reference_golden():
    """oracle/train_port.py reproduces the step of the unmodified reference (tests/golden/train.npz)."""
    g = golden("train")
    x, labels = train_inputs()
    assert sha(x) == str(g["x_sha"]) and np.array_equal(labels, g["labels"])
    assert float(g["pool_tie_margin"]) > 1e-5          # fixture chosen without near-tied max-pool windows
    sd = synth.make_weights(int(g["weight_seed"]))
    port = load_numpy_state(ClassifierPort(31), sd)
    keep = torch.from_numpy(golden_keep(g))
    torch.set_num_threads(4)
    loss, logits, grads = train_port.loss_and_grads(port, torch.from_numpy(x), torch.from_numpy(labels), keep)
    assert abs(loss - float(g["loss"])) < 1e-4
    assert np.max(np.abs(logits.numpy() - g["logits"])) < 1e-4
    pos = sample_positions(sd)
    for k, gr in grads.items():
        flat = gr.numpy().reshape(-1).astype(np.float64)
        if k == "attention.bias":
            assert abs(flat[0]) < 1e-5
            continue
        scale = max(np.max(np.abs(g[f"gsamp/{k}"])), float(g[f"gnorm/{k}"]) / np.sqrt(flat.size))
        assert np.max(np.abs(flat[pos[k]] - g[f"gsamp/{k}"])) < 2e-4 * scale, k
        assert abs(np.sqrt((flat * flat).sum()) - float(g[f"gnorm/{k}"])) < 2e-4 * float(g[f"gnorm/{k}"]), k
    for k, b in port.named_buffers():
        if "num_batches" not in k:
            assert np.max(np.abs(b.numpy() - g[f"buf/{k}"])) < 1e-5, k
    # Adam step of the reference (scripts/train.py:246-250: coupled weight decay)
    opt = torch.optim.Adam(port.parameters(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    opt.step()
    new = dict(port.named_parameters())
    for k in grads:
        if k != "attention.bias":
            assert np.max(np.abs(new[k].detach().numpy().reshape(-1)[pos[k]] - g[f"psamp/{k}"])) < 2e-5, k


def test_loss_scaler_mirrors_gradscaler_bookkeeping():
    s = train.LossScaler(enabled=True, init_scale=1024.0, growth_interval=3)
    for _ in range(3):
        s.update(False)
    assert s.scale == 2048.0
    s.update(True)
    assert s.scale == 1024.0
    s.update(False)
    s.update(False)
    s.update(True)                      # back-off resets the growth streak
    assert s.scale == 512.0
    off = train.LossScaler(enabled=False)
    off.update(True)
    assert off.scale == 1.0


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 30043):
        for world in (1, 2, 3, 8):
            spans = [train.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        # gradients + trailing found-inf flag, as DataParallelTrainer lays them out
        flat = torch.full((n + 1,), float(rank + 1))
        flat[n] = 0.0
        got_world = train.sync_flat_gradients(flat)
        assert got_world == world
        assert torch.all(flat[:n] == 3.0) and flat[n] == 0.0          # 1 + 2, averaged later via inv_scale
        # step 2: only rank 1 sees a non-finite gradient -> every rank must see the flag
        flat = torch.full((n + 1,), 0.5)
        flat[n] = 1.0 if rank == 1 else 0.0
        train.sync_flat_gradients(flat)
        assert flat[n] == 1.0
        # replicas that were initialised differently start from rank 0's state (no silent divergence)
        state = torch.full((50,), float(10 + rank))
        assert train.broadcast_initial_state(state) and torch.all(state == 10.0)
        a, b = train.shard_range(30043, rank, world)
        counts = torch.tensor([b - a])
        dist.all_reduce(counts)
        assert int(counts) == 30043
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gradient_sync_world_size_2_gloo(tmp_path):
    """The N>1 path of the training step on CPU: one all-reduce carries gradients and the collective skip flag."""
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
