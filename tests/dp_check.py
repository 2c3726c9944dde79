"""2-rank NCCL check of the data-parallel training step (launched by tests/test_gpu_train.py under torchrun).

Every rank steps its own shard; afterwards (a) all ranks hold bit-identical parameters, (b) the all-reduced flat
gradient equals the sum of the two ranks' local gradients computed separately on rank 0, (c) a non-finite input on
ONE rank makes BOTH ranks skip the update.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import synth, train_inputs  # noqa: E402

models = importlib.import_module("speech-intent-recognizer_b200.models.models")
train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")


def build():
    sd = synth.make_weights(1234)
    m = models.CNNAudioGRU(31)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    return m.cuda()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = 8
    x, labels = train_inputs(seed=43, batch=B * world)
    keep = (np.random.default_rng(8).random((B * world, 25, 512)) >= 0.5).astype(np.uint8)
    a, b = train.shard_range(B * world, rank, world)
    xs, ls, ks = (torch.from_numpy(v[a:b]).cuda() for v in (x, labels, keep))

    model = build()
    tr = train.DataParallelTrainer(model, lr=1e-3, weight_decay=1e-4, use_amp=True)
    n = model.weight_count()
    tr.step(xs, ls, dropout_keep=ks)
    reduced = model._flat_grad[:n].clone()
    # (a) identical trainable parameters everywhere (BatchNorm running statistics are per GPU: no SyncBN in the reference)
    mine = model._flat.clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    for o, k in model.param_segments():
        assert torch.equal(mine[o:o + k], other[o:o + k]), "parameters diverged between ranks"
    # (b) the all-reduce is the sum of the local gradients
    if rank == 0:
        total = torch.zeros(n, device="cuda")
        for r in range(world):
            a2, b2 = train.shard_range(B * world, r, world)
            solo = build()
            solo.train()
            solo._next_dropout_keep = torch.from_numpy(keep[a2:b2]).cuda()
            out = solo(torch.from_numpy(x[a2:b2]).cuda())
            loss = torch.nn.functional.cross_entropy(out, torch.from_numpy(labels[a2:b2]).cuda())
            loss.backward()
            total += solo._flat_grad[:n]
        err = float((reduced / tr.scaler.scale - total).abs().max() / total.abs().max())
        assert err < 1e-5, err
    # (c) collective skip
    before = model._flat.clone()
    bad = xs.clone()
    if rank == 1:
        bad[0, 0, 0] = float("inf")
    tr.step(bad, ls, dropout_keep=ks)
    assert tr.skipped_steps == 1 and tr.adam_steps == 1
    for o, k in model.param_segments():
        assert torch.equal(model._flat[o:o + k], before[o:o + k])
    # (d) replicas built from DIFFERENT seeds (the reference never seeds its init): the trainer's constructor broadcasts
    # rank 0's parameters and BatchNorm buffers, so the ranks start - and after a step remain - identical
    torch.manual_seed(100 + rank)
    m2 = models.CNNAudioGRU(31).cuda()
    tr2 = train.DataParallelTrainer(m2, lr=1e-3, weight_decay=1e-4, use_amp=True)
    for when in ("after construction", "after one step"):
        mine = m2._flat.clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        if when == "after construction":
            assert torch.equal(mine, other), "initial state (parameters + buffers) not broadcast"
            tr2.step(xs, ls, dropout_keep=ks)
        else:
            for o, k in m2.param_segments():
                assert torch.equal(mine[o:o + k], other[o:o + k]), "differently seeded replicas diverged"
    # (e) the graph-replayed step (backward split in two, the GRU / head bucket all-reduced on a side stream under the conv
    # backward, the conv bucket after it) against the eager step with its single all-reduce: same losses, same parameters,
    # same collective skip
    me, mg = build(), build()
    te = train.DataParallelTrainer(me, lr=1e-3, weight_decay=1e-4, use_amp=True, seed=3)
    tg = train.DataParallelTrainer(mg, lr=1e-3, weight_decay=1e-4, use_amp=True, seed=3, use_graph=True)
    for i in range(4):
        inp = bad if i == 2 else xs
        le, lg = te.step(inp, ls), tg.step(inp, ls)
        assert tg._graph is not None and "mid" in tg._graph, "the bucketed graph path did not run"
        assert (le == lg) or (le != le and lg != lg), (i, le, lg)
        for o, k in me.param_segments():
            assert torch.equal(me._flat[o:o + k], mg._flat[o:o + k]), f"graph step {i} differs from the eager step"
    assert te.skipped_steps == 1 and tg.skipped_steps == 1 and tg.adam_steps == 3
    mine = mg._flat.clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    for o, k in mg.param_segments():
        assert torch.equal(mine[o:o + k], other[o:o + k]), "graph-replayed replicas diverged"
    dist.barrier()
    if rank == 0:
        print("dp_check ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
