"""GPU parity tests of the training step (through the C ABI) against the golden step produced by the reference
itself (tests/golden/train.npz) and against the oracle restatement (oracle/train_port.py) at config-4 batch size.

Bars: logits within 1e-3 absolute; gradients within 1e-3 of each tensor's scale (fp32 everywhere, different
summation orders); BatchNorm running statistics within 1e-5; Adam-updated parameters within 2 % of the step size.
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import train_port
from oracle.torch_port import ClassifierPort, load_numpy_state
from tests.util import (LOGIT_ABS_TOL, ROOT, TRAIN_SEED_B16, golden, golden_keep, pool_tie_margin, sample_positions, synth,
                        train_inputs)

pytestmark = pytest.mark.gpu
native = importlib.import_module("speech-intent-recognizer_b200._native")
models = importlib.import_module("speech-intent-recognizer_b200.models.models")
train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")

GRAD_REL_TOL = 1e-3


def make_model(seed=1234):
    sd = synth.make_weights(seed)
    m = models.CNNAudioGRU(31)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    return m.cuda(), sd


def check_grads_against_golden(g, sd, grads_by_key):
    pos = sample_positions(sd)
    for k, gr in grads_by_key.items():
        flat = gr.detach().cpu().numpy().reshape(-1).astype(np.float64)
        if k == "attention.bias":
            assert abs(flat[0]) < 1e-4
            continue
        rms = float(g[f"gnorm/{k}"]) / np.sqrt(flat.size)
        scale = max(np.max(np.abs(g[f"gsamp/{k}"])), rms)
        assert np.max(np.abs(flat[pos[k]] - g[f"gsamp/{k}"])) < GRAD_REL_TOL * scale, k
        assert abs(np.sqrt((flat * flat).sum()) - float(g[f"gnorm/{k}"])) < GRAD_REL_TOL * float(g[f"gnorm/{k}"]), k
        assert abs(flat.sum() - float(g[f"gsum/{k}"])) < GRAD_REL_TOL * float(g[f"gnorm/{k}"]) * np.sqrt(flat.size), k


def check_params_against_golden(g, sd, model, lr):
    pos = sample_positions(sd)
    new = {k: v.detach().cpu().numpy().reshape(-1) for k, v in model.named_parameters()}
    for k in new:
        if k == "attention.bias":
            continue
        gs = np.abs(g[f"gsamp/{k}"])
        solid = gs > 1e-3 * gs.max()                      # Adam's first step is lr * g / (|g| + eps): ill-conditioned at g ~ 0
        diff = np.abs(new[k][pos[k]] - g[f"psamp/{k}"])
        assert np.all(diff[solid] < 0.02 * lr), (k, diff[solid].max())
        assert np.all(diff < 2.1 * lr), k


def test_train_forward_backward_match_reference_golden():
    g = golden("train")
    x, labels = train_inputs()
    model, sd = make_model(int(g["weight_seed"]))
    model.train()
    model._next_dropout_keep = torch.from_numpy(golden_keep(g)).cuda()
    before = native.launch_count()
    out = model(torch.from_numpy(x).cuda())
    assert np.max(np.abs(out.detach().cpu().numpy() - g["logits"])) < LOGIT_ABS_TOL
    loss = torch.nn.functional.cross_entropy(out, torch.from_numpy(labels).cuda())
    assert abs(float(loss) - float(g["loss"])) < 1e-3
    loss.backward()
    assert native.launch_count() - before > 30
    check_grads_against_golden(g, sd, {k: p.grad for k, p in model.named_parameters()})
    for k, b in model.named_buffers():                                  # running statistics, momentum 0.1, unbiased var
        if "num_batches" in k:
            assert int(b) == 1
        else:
            assert np.max(np.abs(b.cpu().numpy() - g[f"buf/{k}"])) < 1e-5, k
    # the reference's optimizer on top of our gradients (scripts/train.py:246-250, :107-108)
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    opt.step()
    check_params_against_golden(g, sd, model, float(g["lr"]))
    # eval after a training step re-folds BatchNorm from the updated running statistics
    model.eval()
    ref = load_numpy_state(ClassifierPort(31), {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}).eval()
    with torch.no_grad():
        want = ref(torch.from_numpy(x))
    assert np.max(np.abs(model(torch.from_numpy(x).cuda()).cpu().numpy() - want.numpy())) < LOGIT_ABS_TOL


def test_fused_trainer_step_matches_reference_golden():
    """DataParallelTrainer (fused CE, loss scaling, unscale + Adam) on one GPU == the reference's step."""
    g = golden("train")
    x, labels = train_inputs()
    model, sd = make_model(int(g["weight_seed"]))
    tr = train.DataParallelTrainer(model, lr=float(g["lr"]), weight_decay=float(g["weight_decay"]), use_amp=True)
    keep = torch.from_numpy(golden_keep(g)).cuda()
    loss = tr.step(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), dropout_keep=keep)
    assert abs(loss - float(g["loss"])) < 1e-3
    assert tr.adam_steps == 1 and tr.skipped_steps == 0 and tr.scaler.scale == 65536.0
    flat_grad = model._flat_grad[:model.weight_count()] / 65536.0
    grads, off = {}, 0
    named = dict(model._named_tensors())
    for key, shape in model._spec():
        n = int(np.prod(shape))
        if key in dict(model.named_parameters()):
            grads[key] = flat_grad[off:off + n]
        off += n
    check_grads_against_golden(g, sd, grads)
    check_params_against_golden(g, sd, model, float(g["lr"]))
    assert np.max(np.abs(named["bn2.running_var"].cpu().numpy() - g["buf/bn2.running_var"])) < 1e-5
    # a non-finite gradient skips the update on this rank and backs the scale off
    before = model._flat.clone()
    bad = torch.from_numpy(x).cuda()
    bad[0, 0, 0] = float("inf")
    tr.step(bad, torch.from_numpy(labels).cuda(), dropout_keep=keep)
    assert tr.skipped_steps == 1 and tr.adam_steps == 1 and tr.scaler.scale == 32768.0
    segs = model.param_segments()
    for o, n in segs:
        assert torch.equal(model._flat[o:o + n], before[o:o + n])


def test_config4_batch_gradients_match_oracle():
    """Batch 16 x [64,200] (config.yaml batch_size): every gradient element against the CPU restatement."""
    B = 16
    x, labels = train_inputs(seed=TRAIN_SEED_B16, batch=B)
    rng = np.random.default_rng(5)
    keep = (rng.random((B, 25, 512)) >= 0.5).astype(np.uint8)
    model, sd = make_model(1234)
    model.train()
    model._next_dropout_keep = torch.from_numpy(keep).cuda()
    out = model(torch.from_numpy(x).cuda())
    loss = torch.nn.functional.cross_entropy(out, torch.from_numpy(labels).cuda())
    loss.backward()
    port = load_numpy_state(ClassifierPort(31), sd)
    torch.set_num_threads(os.cpu_count() or 1)
    # the fixture has no near-tied max-pool window, so arg-max routing is implementation independent
    assert pool_tie_margin(port, torch.from_numpy(x)) > 2e-6
    want_loss, want_logits, want = train_port.loss_and_grads(port, torch.from_numpy(x), torch.from_numpy(labels),
                                                            torch.from_numpy(keep))
    assert np.max(np.abs(out.detach().cpu().numpy() - want_logits.numpy())) < LOGIT_ABS_TOL
    assert abs(float(loss) - want_loss) < 1e-3
    for k, p in model.named_parameters():
        a, b = p.grad.cpu().numpy().astype(np.float64), want[k].numpy().astype(np.float64)
        if k == "attention.bias":
            assert abs(a[0]) < 1e-4
            continue
        assert np.max(np.abs(a - b)) < GRAD_REL_TOL * np.max(np.abs(b)), (k, np.max(np.abs(a - b)), np.max(np.abs(b)))


def test_cross_entropy_and_adam_kernels_against_torch():
    torch.manual_seed(3)
    logits = (torch.randn(37, 31, device="cuda") * 6).requires_grad_()
    labels = torch.randint(0, 31, (37,), device="cuda")
    loss, dl = native.cross_entropy(logits.detach(), labels, scale=8.0)
    want = torch.nn.functional.cross_entropy(logits, labels)
    want.backward()
    assert abs(float(loss) - float(want)) < 1e-5
    assert torch.allclose(dl / 8.0, logits.grad, atol=1e-7, rtol=1e-5)
    # Adam over two segments with a gap (the gap = BatchNorm running statistics must not move), 3 steps
    n = 5000
    p0 = torch.randn(n, device="cuda")
    ref_p = p0.clone().requires_grad_()
    opt = torch.optim.Adam([ref_p], lr=3e-3, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    segs = [(0, 2000), (2100, 2900)]
    for step in range(1, 4):
        gr = torch.randn(n, device="cuda")
        ref_p.grad = gr.clone()
        opt.step()
        native.adam_step(p, gr * 4.0, m, v, segs, 3e-3, (0.9, 0.999), 1e-8, 1e-2, step=step, inv_scale=0.25)
    assert torch.allclose(p[:2000], ref_p.detach()[:2000], atol=1e-6, rtol=1e-5)
    assert torch.allclose(p[2100:], ref_p.detach()[2100:], atol=1e-6, rtol=1e-5)
    assert torch.equal(p[2000:2100], p0[2000:2100])
    flag = torch.ones(1, device="cuda")
    snap = p.clone()
    native.adam_step(p, gr, m, v, segs, 3e-3, step=4, found_inf=flag)
    assert torch.equal(p, snap)
    flag.zero_()
    gr[123] = float("nan")
    native.grad_nonfinite(gr, n, flag)
    assert float(flag) == 1.0


def test_train_epoch_mirror_runs_and_learns():
    """scripts/train.py:72-118 mirror with torch's optimizer / criterion on top of the CUDA model."""
    model, _ = make_model(1234)
    x, labels = train_inputs(seed=9, batch=8)
    loader = [(torch.from_numpy(x), torch.from_numpy(labels))] * 6
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    first = train.train_epoch(model, loader[:1], opt, crit, "cuda")
    for _ in range(3):
        last = train.train_epoch(model, loader, opt, crit, "cuda")
    assert last < 0.85 * first
    vloss, acc = train.validate(model, loader[:1], crit, "cuda")
    assert np.isfinite(vloss) and 0.0 <= acc <= 1.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_two_gpus_equals_one_big_batch_of_gradients():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "dp_check ok" in out.stdout


def test_graph_replayed_step_equals_eager_step():
    """``DataParallelTrainer(use_graph=True)``: the step captured once as two CUDA graphs (device-resident step state: Adam
    step count / bias corrections, loss scale, dropout stream position) and replayed == the eager step over several batches,
    through a skipped (non-finite) step and a loss-scale back-off."""
    x, labels = train_inputs(seed=21, batch=16 * 5)
    xs = [torch.from_numpy(x[16 * i:16 * (i + 1)]).cuda() for i in range(5)]
    ys = [torch.from_numpy(labels[16 * i:16 * (i + 1)]).cuda() for i in range(5)]
    results = []
    for use_graph in (False, True):
        model, _ = make_model(1234)
        tr = train.DataParallelTrainer(model, lr=1e-3, weight_decay=1e-4, use_amp=True, seed=3, use_graph=use_graph)
        losses = [tr.step(xs[i], ys[i]) for i in range(3)]
        bad = xs[3].clone()
        bad[0, 0, 0] = float("inf")
        before = model._flat.clone()
        losses.append(tr.step(bad, ys[3]))                                  # every rank skips; the scale backs off
        for o, k in model.param_segments():
            assert torch.equal(model._flat[o:o + k], before[o:o + k]), "a skipped step changed parameters"
        assert tr.skipped_steps == 1 and tr.adam_steps == 3 and tr.scaler.scale == 32768.0
        losses.append(tr.step(xs[4], ys[4]))
        if use_graph:
            assert tr._graph is not None and tr.graph_replays == 5 and tr._graph["launches"] > 20
            step_d, scale_d, off_d = tr._graph["state"].read()
            assert step_d == 4 and scale_d == 32768.0 and off_d == model._dropout_offset
        keep = torch.cat([torch.arange(o, o + k) for o, k in model.param_segments()]).cuda()   # trainable parameters only: the
        # BatchNorm running statistics took the non-finite batch in (train-mode forward), in both runs alike
        results.append((losses, model._flat[keep].clone(), tr.exp_avg[keep].clone(), tr.exp_avg_sq[keep].clone()))
    (l0, p0, m0, v0), (l1, p1, m1, v1) = results
    # same kernels in the same order; the only difference is WHERE the Adam bias corrections are evaluated (host libm pow vs
    # device pow, both in double): the last bit of 1 - beta^step may differ, i.e. ~1e-7 relative on the update
    assert np.allclose(l0[:3], l1[:3], rtol=1e-5) and np.isfinite(l0[4]) and abs(l0[4] - l1[4]) < 1e-4 * abs(l0[4])
    assert torch.equal(m0, m1) or torch.allclose(m0, m1, rtol=1e-5, atol=1e-9)
    assert torch.allclose(v0, v1, rtol=1e-5, atol=1e-12)
    assert float((p0 - p1).abs().max()) < 1e-6


def test_backward_in_two_parts_equals_one_call():
    """``sir_model_backward_part`` 1 (head + GRU layers) then 2 (conv stack) writes the same bits as ``sir_model_backward``;
    part 1 alone leaves the conv gradients zero and already holds every gradient behind ``sir_model_gru_grad_offset``."""
    x, labels = train_inputs(seed=TRAIN_SEED_B16, batch=16)
    model, _ = make_model(1234)
    model._ensure_native()
    flat, nm, n = model.flatten_parameters_(), model._native_model, model.weight_count()
    off = nm.gru_grad_offset()
    assert 0 < off < n
    feats = torch.from_numpy(x).cuda()
    grads = []
    for split in (False, True):
        p = flat.clone()                                               # (the forward updates the running statistics in place)
        logits = nm.train_forward(p, feats, seed=7, offset=0, bn_momentum=0.1, bn_eps=1e-5)
        _, dlogits = native.cross_entropy(logits, torch.from_numpy(labels).cuda(), scale=1024.0)
        g = torch.full((n + 4,), 3.0, device="cuda")
        if split:
            nm.backward_part(p, dlogits, g, 1)
            assert float(g[:off].abs().max()) == 0.0 and float(g[off:n].abs().max()) > 0.0
            tail = g[off:n].clone()
            nm.backward_part(p, None, g, 2)
            assert torch.equal(g[off:n], tail), "part 2 touched the first bucket"
        else:
            nm.backward(p, dlogits, g)
        grads.append(g[:n].clone())
    assert torch.equal(grads[0], grads[1])
    assert float(grads[0][:off].abs().max()) > 0.0
    with pytest.raises(native.NativeError):
        nm.backward_part(p, dlogits, g, 3)
