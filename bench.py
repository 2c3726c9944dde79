#!/usr/bin/env python
"""bench.py - utterances/sec of the B200-native hot path, with roofline, CPU baseline and the end-to-end number.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|config4|config5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`; the default is the one the metric is quoted on):
  config2 (default)  synthetic FSC-shaped inference batch: 256 utterances x 3 s @16 kHz per GPU, config.yaml model (64
                     mels, 200 frames, 31 classes).  A step = fused log-mel frontend + CNNAudioGRU forward over one batch.
                     The line also carries a `train` object: the config-4 training step (batch 16 per GPU) with its
                     gradient all-reduce at N > 1, so the driver's scaling run exercises the collective.
  config3            feature precompute: 30,043 utterances x 3 s sharded contiguously over the ranks, frontend only; a step
                     = one pass over the rank's shard.
  config4            data-parallel training: batch 16 per GPU, SpecAugment from the device sampler on an HBM-resident
                     feature cache, fused step, ONE NCCL all-reduce of the flat gradients per step.
  config5            long audio: 512 utterances x 10 s per GPU (4096 over 8), 80-mel frontend + the 80-mel classifier.
Utterances are independent, so N GPUs run N shards with no data-path collective ("scaling": "weak"); only config4 (and the
`train` object) has an exchange step.

 value  : whole-job utterances/s with inputs resident in HBM, K back-to-back steps between CUDA events, max over ranks.
          Inputs rotate over buffers larger than the 126 MB L2 (or are larger than L2 themselves).
 e2e    : the same metric through the public Python API from HOST buffers holding 16-bit PCM - what a WAV file holds and
          what the reference's loader starts from (scripts/precompute_features.py:47) - H2D copy + pipeline + D2H of the
          result inside the timed region, every step.  `e2e_fp32` is the same loop from fp32 host buffers.
 roofline / frontend_roofline / slowest_stage_roofline / rooflines / stages : per-kernel device times from a separate K-step
          pass with CUDA events around each stage on the launching stream (sir_profile_*), algorithmic flops / bytes per
          DESIGN.md section 4.  `roofline` is the frontend's (the kernel BASELINE.json's metric names, HBM-bound);
          `slowest_stage_roofline` the slowest single stage's (the GRU recurrence: a latency chain), `rooflines` all of them.
 cpu_baseline : the reference's CPU path on this box's host cores (rank 0, N = 1 only): the reference's own classes from
          baseline/_ref when build() could copy them ("reference"), else oracle/torch_port.py ("port").
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_CLASSES, OUT_FRAMES = 31, 200
WORKLOADS = {
    "config2": {"batch": 256, "samples": 48000, "n_mels": 64, "steps": 50, "max_duration": 5.0,
                "name": "configs[1]: 256 utt x 3 s @16 kHz per GPU, config.yaml model (64 mel, 200 frames, 31 classes), "
                        "seeded random-init weights"},
    "config3": {"total": 30043, "samples": 48000, "n_mels": 64, "steps": 10,
                "name": "configs[2]: feature precompute of 30,043 synthetic utt x 3 s @16 kHz, sharded contiguously over "
                        "the ranks, frontend only (unpadded [64, 94] features)"},
    "config4": {"batch": 16, "samples": 48000, "n_mels": 64, "steps": 100,
                "name": "configs[3]: data-parallel training, batch 16 per GPU (config.yaml), SpecAugment (augment_prob 0.7, "
                        "masks 20/10) on an HBM-resident synthetic feature cache, Adam lr 5e-5 wd 1e-4, loss scaling, one "
                        "NCCL gradient all-reduce per step"},
    "config5": {"batch": 512, "samples": 160000, "n_mels": 80, "steps": 20, "max_duration": None,
                "name": "configs[4]: 512 utt x 10 s @16 kHz per GPU (4096 over 8), 80-mel frontend (313 frames normalised, "
                        "trimmed to 200) + the 80-mel classifier (gru_input_size 1280)"},
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the config2 workload
# (256 utterances): profiles/ncu_traffic_b256.json, which names the capture each number comes from.
NCU_TRAFFIC_B256 = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_b256.json"))) \
    if os.path.exists(os.path.join(ROOT, "profiles", "ncu_traffic_b256.json")) else {}
SPLIT_PASSES = {"conv2_bn_relu_pool": 3, "conv3_bn_relu_pool": 3, "gru_l0_input_gemm": 3, "gru_l1_input_gemm": 3,
                "gru_l0_recurrence": 3, "gru_l1_recurrence": 3}


def flops_per_utt(n_mels):
    """SURVEY.md 8 a11 (2 flops per MAC) at [n_mels, 200]; the conv stack and the layer-0 projection scale with n_mels."""
    s = n_mels / 64.0
    return {"conv1_bn_relu_pool": 7.37e6 * s, "conv2_bn_relu_pool": 117.96e6 * s, "conv3_bn_relu_pool": 117.96e6 * s,
            "gru_l0_input_gemm": 78.64e6 * s, "gru_l0_recurrence": 19.66e6, "gru_l1_input_gemm": 39.32e6,
            "gru_l1_recurrence": 19.66e6, "attention_fc": 0.06e6}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (hbm copy; bf16 sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback of B200_PROFILING.md"}


class ClockSampler:
    """SM clock and throttle reasons sampled in-process through NVML (pynvml) every few ms, so that even a
    20 ms timed region gets samples; `mark()` brackets the timed region, samples inside it are reported."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index, period_s=0.002):
        self.index, self.period, self.rows, self.marks = index, period_s, [], []
        self.handle = self.thread = None
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001 - NVML missing: report it, never fail the bench
            self.error = f"nvml unavailable: {e}"

    def _sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:  # noqa: BLE001
            reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append((time.perf_counter(), sm, reasons))

    def _pump(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def start(self):
        if self.handle is not None:
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()

    def mark(self):
        """Call at the start and end of every timed region (host time; the regions are device-synchronised)."""
        if self.handle is not None:
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                pass
        self.marks.append(time.perf_counter())

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [getattr(self, "error", "nvml unavailable")],
                    "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=1.0)
        spans = list(zip(self.marks[0::2], self.marks[1::2]))
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in spans)] or self.rows
        sm = [r[1] for r in inside]
        bits = 0
        for r in inside:
            bits |= r[2]
        reasons = sorted(n for n, m in self.REASONS if bits & m)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "NVML in-process, samples inside the timed regions (value + e2e)"}


def synth_batch(native_synth, seed, batch, samples):
    """Speech-like rows are expensive to synthesise on the host; tile 32 distinct utterances with per-row gains."""
    base = native_synth.speech_like(seed, 32, samples)
    reps = (batch + 31) // 32
    gains = np.linspace(0.25, 1.0, reps * 32, dtype=np.float32)[:, None]
    return (np.tile(base, (reps, 1)) * gains)[:batch]


def to_pcm16(x: torch.Tensor) -> torch.Tensor:
    return (x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16)


# ------------------------------------------------------------------------------------------------------------------
# the reference's CPU path (test / bench infrastructure only - never on the product path)
# ------------------------------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's CPU implementation of the path on the host cores, starting - like the reference's loader - from
    16-bit PCM: int16 -> float / 32768 (what torchaudio.load returns, scripts/precompute_features.py:47), then per
    utterance ``mel_transform -> amplitude_to_db -> normalise`` in a Python loop like :124-130, pad to 200
    (scripts/dataset.py:109-113), then ``CNNAudioGRU.eval()`` fp32 forward in one batch (scripts/evaluate.py:79-83).

    kind "reference": the reference's own ``AudioFeatureExtractor`` / ``CNNAudioGRU`` classes, imported unmodified from
    baseline/_ref (copied there by __graft_entry__.build() in the build container; git-ignored, travels with the
    snapshot).  kind "port": oracle/torch_port.py, which issues the same torchaudio / torch.nn calls."""

    def __init__(self, native_synth, n_utts, samples, n_mels=64, classifier=True, threads=None, max_duration=5.0):
        from oracle.torch_port import ClassifierPort, FeaturePort, load_numpy_state
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.n, self.n_mels, self.max_duration = n_utts, n_mels, max_duration
        self.pcm = to_pcm16(torch.from_numpy(synth_batch(native_synth, 99, n_utts, samples)))
        self.kind, self.ref_extractor, self.model = "port", None, None
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        sd = native_synth.make_weights(1234, NUM_CLASSES, n_mels) if classifier else None
        if n_mels == 64 and os.path.exists(os.path.join(ref_dir, "scripts", "precompute_features.py")):
            try:
                sys.path.insert(0, ref_dir)
                ref_pre = importlib.import_module("scripts.precompute_features")
                ref_models = importlib.import_module("models.models")
                self.ref_extractor = ref_pre.AudioFeatureExtractor()
                if classifier:
                    self.model = load_numpy_state(ref_models.CNNAudioGRU(NUM_CLASSES).eval(), sd)
                self.kind = "reference"
            except Exception as e:  # noqa: BLE001 - fall back to the port, say why
                self.ref_error = f"{type(e).__name__}: {e}"
                self.ref_extractor = None
            finally:
                sys.path.remove(ref_dir)
        if self.ref_extractor is None:
            self.fp = FeaturePort(n_mels=n_mels)
            if classifier:
                self.model = load_numpy_state(ClassifierPort(NUM_CLASSES, n_mels=n_mels).eval(), sd)
        torch.set_num_threads(self.threads)

    def features(self):
        out = []
        limit = int(self.max_duration * 16000) if self.max_duration is not None else None
        for i in range(self.n):
            w = self.pcm[i:i + 1].to(torch.float32) / 32768.0                     # the loader's int16 -> float
            if self.ref_extractor is not None:                                   # scripts/precompute_features.py:59-73
                ex = self.ref_extractor
                if limit is not None and w.shape[1] > limit:
                    w = w[:, :limit]
                m = ex.amplitude_to_db(ex.mel_transform(w)).squeeze(0)
                m = (m - m.mean()) / (m.std() + 1e-5)
            else:
                m = self.fp.one(w, self.max_duration)
            if m.shape[1] > OUT_FRAMES:                                          # scripts/dataset.py:109-113
                m = m[:, :OUT_FRAMES]
            elif m.shape[1] < OUT_FRAMES:
                m = torch.nn.functional.pad(m, (0, OUT_FRAMES - m.shape[1]))
            out.append(m)
        return torch.stack(out)

    def step(self):
        """One pass over the n_utts batch -> (feature seconds, forward seconds)."""
        t0 = time.perf_counter()
        feats = self.features()
        t1 = time.perf_counter()
        if self.model is not None:
            with torch.no_grad():
                self.model(feats)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def run_for(self, min_seconds, min_steps=2, max_steps=200):
        feat_s = fwd_s = 0.0
        steps = 0
        while steps < min_steps or (feat_s + fwd_s < min_seconds and steps < max_steps):
            a, b = self.step()
            feat_s, fwd_s, steps = feat_s + a, fwd_s + b, steps + 1
        return steps, feat_s, fwd_s

    def describe(self, steps, feat_s, fwd_s):
        what = ("the reference's AudioFeatureExtractor / CNNAudioGRU classes from baseline/_ref (file decode replaced by "
                "in-memory int16 -> float)" if self.kind == "reference" else "oracle/torch_port.py (same torchaudio / torch.nn calls)")
        return (f"{steps} passes over {self.n} utterances ({feat_s + fwd_s:.1f} s of CPU work), {what}: int16 -> float + "
                f"per-utterance MelSpectrogram+AmplitudeToDB+normalise loop {feat_s / steps:.3f} s"
                + (f" + CNNAudioGRU fp32 batched forward {fwd_s / steps:.3f} s" if self.model is not None else "") + " per pass")


class CpuTrainArm:
    """The reference's training step on the host cores in fp32 (scripts/train.py:80-116 without autocast - the CPU path has
    none): zero_grad, forward in train mode (nn.GRU's own dropout), CrossEntropyLoss, backward, Adam, loss.item()."""

    def __init__(self, native_synth, batch, threads=None):
        from oracle.torch_port import ClassifierPort, load_numpy_state
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.model = load_numpy_state(ClassifierPort(NUM_CLASSES).train(), native_synth.make_weights(1234))
        self.opt = torch.optim.Adam(self.model.parameters(), lr=5e-5, weight_decay=1e-4)
        self.x, self.y = train_features(7, batch)
        self.batch = batch

    def step(self):
        self.opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(self.model(self.x), self.y)
        loss.backward()
        self.opt.step()
        return loss.item()


def train_features(seed, n, frames=OUT_FRAMES, valid=94, n_mels=64):
    """Normalised log-mel-like feature maps [n, n_mels, frames] with a zero tail beyond `valid` frames + uniform labels."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, n_mels, frames), generator=g)
    x += 0.8 * torch.sin(torch.linspace(0, 6, n_mels))[None, :, None]
    x[:, :, valid:] = 0.0
    y = torch.randint(0, NUM_CLASSES, (n,), generator=g)
    return x, y


class JsonStdout:
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on
    init), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to the saved fd."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.fd, (json.dumps(obj) + "\n").encode())


# ------------------------------------------------------------------------------------------------------------------
# --impl reference
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, out):
    """The reference's own CPU implementation of the path, all host threads, rank 0 only; one step = a bounded sample of
    the workload (config2: the full 256-utterance batch)."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    native_synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
    t_start = time.perf_counter()
    if args.workload == "config4":
        arm = CpuTrainArm(native_synth, wl["batch"])
        for _ in range(min(args.warmup, 2)):
            arm.step()
        t0 = time.perf_counter()
        n_steps = min(args.steps, 20)
        for _ in range(n_steps):
            arm.step()
        total = time.perf_counter() - t0
        n_per_step, threads, kind = wl["batch"], arm.threads, "port"
        sample = (f"{n_steps} training steps of batch {wl['batch']} (fp32: zero_grad, train-mode forward, CrossEntropyLoss, "
                  f"backward, Adam, loss.item()) through oracle/torch_port.ClassifierPort = the reference's torch.nn calls")
    else:
        n_per_step = {"config2": wl.get("batch"), "config3": 256, "config5": 16}[args.workload]
        arm = CpuArm(native_synth, n_per_step, wl["samples"], wl["n_mels"], classifier=args.workload != "config3",
                     max_duration=wl.get("max_duration", 5.0))
        for _ in range(min(args.warmup, 3)):
            arm.step()
        feat_s = fwd_s = 0.0
        n_steps = args.steps
        for _ in range(n_steps):
            a, b = arm.step()
            feat_s, fwd_s = feat_s + a, fwd_s + b
        total = feat_s + fwd_s
        threads, kind = arm.threads, arm.kind
        sample = arm.describe(n_steps, feat_s, fwd_s) + f"; wall {time.perf_counter() - t_start:.1f} s"
    value = n_per_step * n_steps / total
    out.emit({
        "impl": "reference", "metric": "utterances/sec (features+forward)", "value": value, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": n_steps, "warmup": args.warmup, "ms_per_step": total / n_steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "batch_per_step": n_per_step},
        "cpu_baseline": {"value": value, "unit": "utt/s", "cores": threads, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count() or 1, "torch": torch.__version__},
        "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------------
# shared pieces of the B200 arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.native = importlib.import_module("speech-intent-recognizer_b200._native")
        self.synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
        self.pre = importlib.import_module("speech-intent-recognizer_b200.scripts.precompute_features")
        self.models = importlib.import_module("speech-intent-recognizer_b200.models.models")
        self.pipe_mod = importlib.import_module("speech-intent-recognizer_b200.pipeline")
        self.all_cpus = os.sched_getaffinity(0)
        self.bound_cpus = self.pipe_mod.bind_host_to_gpu(self.local) if args.numa_bind else None
        self.stream = torch.cuda.current_stream()
        self.sampler = ClockSampler(self.local)
        self.peaks = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(self, x):
        if self.world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, x):
        if self.world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, mark=True):
        """Device time (ms, CUDA events on the launching stream) of fn() between barriers; marks the clock sampler."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if mark:
            self.sampler.mark()
        e0.record(self.stream)
        res = fn()
        e1.record(self.stream)
        self.barrier()
        if mark:
            self.sampler.mark()
        return e0.elapsed_time(e1), res

    def host_binding(self):
        return (f"rank pinned to the {len(self.bound_cpus)} CPU cores NVML lists for its GPU before allocating pinned buffers"
                if self.bound_cpus else "none")

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def stable_repeats(run, repeats, warm_min=1, warm_max=8, tol=0.03, agree=None):
    """Warm `run()` (a wall-clock timed K-step loop returning seconds) until two consecutive runs agree within `tol`, then
    take `repeats` measured runs.  Returns (median seconds, [measured seconds], warm runs used).
    `agree` (ctx.reduce_max) makes the stop decision COLLECTIVE: `run` holds barriers, so every rank must warm the same
    number of times - with a per-rank decision one rank can leave the loop a run earlier than the others and the job hangs
    in the next barrier (seen at 8 GPUs on the config-3 workload)."""
    prev, used = None, 0
    for used in range(1, warm_max + 1):
        t = run()
        if agree is not None:
            t = agree(t)
        if prev is not None and used > warm_min and abs(t - prev) <= tol * max(t, prev):
            break
        prev = t
    times = [run() for _ in range(max(1, repeats))]
    return float(np.median(times)), times, used


def h2d_rate(ctx, host, reps=10):
    """Pinned host -> device copy time (ms) of one buffer, all ranks copying at once."""
    d = torch.empty(host.shape, dtype=host.dtype, device="cuda")
    for _ in range(2):
        d.copy_(host, non_blocking=True)
    ctx.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(ctx.stream)
    for _ in range(reps):
        d.copy_(host, non_blocking=True)
    c1.record(ctx.stream)
    ctx.barrier()
    return ctx.reduce_max(c0.elapsed_time(c1) / reps)


def stage_rooflines(ctx, stages, steps, batch, n_mels, fe_bytes_per_utt, use_traffic):
    """`stages` (name -> (ms, calls)) -> (stage table, list of roofline objects)."""
    flops = flops_per_utt(n_mels)
    peaks = ctx.peaks
    stage_out = {}
    for name, (ms, calls) in stages.items():
        per_step = ms / steps
        entry = {"ms_per_step": round(per_step, 5), "launch_groups_per_step": calls // steps}
        if name in flops:
            entry["tflops"] = round(flops[name] * batch / (per_step * 1e-3) / 1e12, 3)
        stage_out[name] = entry
    step_ms = sum(v["ms_per_step"] for v in stage_out.values()) or 1.0
    out = []
    for name, v in stage_out.items():
        traffic = NCU_TRAFFIC_B256.get(name) if use_traffic else None
        base = {"kernel": name, "ms": v["ms_per_step"], "share_of_step": round(v["ms_per_step"] / step_ms, 4),
                "traffic": traffic, "peak_source": peaks["source"]}
        if name == "logmel_frontend_kernel":
            gbs = fe_bytes_per_utt * batch / (v["ms_per_step"] * 1e-3) / 1e9
            out.append({**base, "bound": "hbm", "achieved": round(gbs, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": round(gbs / peaks["hbm_gbs"], 5), "bytes_per_utt": fe_bytes_per_utt})
        elif name == "conv1_bn_relu_pool":
            nbytes = (4 * n_mels * OUT_FRAMES + 4 * 32 * (n_mels // 2) * (OUT_FRAMES // 2)) * batch
            gbs = nbytes / (v["ms_per_step"] * 1e-3) / 1e9
            out.append({**base, "bound": "hbm", "achieved": round(gbs, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": round(gbs / peaks["hbm_gbs"], 5)})
        elif name in flops:
            ach = v.get("tflops", 0.0)
            r = {**base, "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                 "frac": round(ach / peaks["tflops"], 5)}
            if name in SPLIT_PASSES:
                r["tensor_pipe_flops_factor"] = SPLIT_PASSES[name]
                r["note"] = ("fp32-accurate 3-pass fp16 hi/lo split: the tensor pipe executes 3x the algorithmic FLOPs "
                             f"({round(3 * ach, 1)} TFLOP/s of fp16 MMA work)")
            if "recurrence" in name:
                r["note"] = (f"latency chain of {OUT_FRAMES // 8} dependent time steps in one launch: "
                             f"{v['ms_per_step'] * 1e3 / (OUT_FRAMES // 8):.2f} us per step; " + r.get("note", ""))
            out.append(r)
    return stage_out, out


# ------------------------------------------------------------------------------------------------------------------
# training step (config4, and the `train` object of the config2 line)
# ------------------------------------------------------------------------------------------------------------------
def measure_training(ctx, steps, warmup, batch=16, dataset_size=4096, cpu_ref=True):
    """The config-4 step on every rank: SpecAugment parameters from the device sampler + masking / pad on an HBM-resident
    feature cache, train-mode forward, cross-entropy, backward, ONE all-reduce of the flat gradients (world > 1), fused
    unscale + Adam, loss read-back.  Device time over K steps (CUDA events), max over ranks."""
    train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")
    native = ctx.native
    model = ctx.models.CNNAudioGRU(NUM_CLASSES)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in ctx.synth.make_weights(1234).items()}, strict=False)
    model = model.cuda()
    trainer = train.DataParallelTrainer(model, lr=5e-5, weight_decay=1e-4, use_amp=True, seed=1, use_graph=not ctx.args.no_graph)
    n = dataset_size
    feats, labels = train_features(100 + ctx.rank, n)
    feats, labels = feats.cuda(), labels.cuda()
    frames = torch.full((n,), 94, dtype=torch.int32, device="cuda")
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5 + ctx.rank)).cuda()
    state = {"i": 0, "epoch": 0, "masks": None}

    def next_batch():
        if state["i"] + batch > n or state["masks"] is None:
            state["i"], state["epoch"] = 0, state["epoch"] + 1
            state["masks"] = native.specaugment_sample(1234, state["epoch"] * n, n, 64, 0, frames=frames, augment_prob=0.7)
        idx = perm[state["i"]:state["i"] + batch]
        state["i"] += batch
        x = native.features_finalize(feats.index_select(0, idx), OUT_FRAMES, frames=frames.index_select(0, idx),
                                     masks=state["masks"].index_select(0, idx).contiguous())
        return x, labels.index_select(0, idx)

    def run(k):
        loss = 0.0
        for _ in range(k):
            x, y = next_batch()
            loss = trainer.step(x, y)
        return loss

    run(max(warmup, 3))
    trainer.time_collective = True
    trainer.collective_ms.clear()
    l0 = native.launch_count()
    ms, loss = ctx.timed(lambda: run(steps))
    launches = native.launch_count() - l0
    if trainer._graph is not None:                                   # replayed launches are not seen by the library's counter
        launches += trainer._graph["launches"] * steps
    coll = float(np.mean(trainer.collective_ms)) if trainer.collective_ms else 0.0
    trainer.time_collective = False
    # host-buffer variant: the batch (features + labels) comes from pinned host memory every step, the loss goes back
    host_x, host_y = feats[:batch].cpu().pin_memory(), labels[:batch].cpu().pin_memory()

    def run_host(k):
        t0 = time.perf_counter()
        for _ in range(k):
            trainer.step(host_x.cuda(non_blocking=True), host_y.cuda(non_blocking=True))
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run_host(5)
    ctx.barrier()
    e2e_s = run_host(steps)
    ms, coll, e2e_s = ctx.reduce_max(ms), ctx.reduce_max(coll), ctx.reduce_max(e2e_s)
    native.profile_enable(True)
    run(min(steps, 20))
    stages = native.profile_read()
    native.profile_enable(False)
    k_prof = min(steps, 20)
    out = {
        "ms_per_step": ms / steps, "value": batch * ctx.world * steps / (ms * 1e-3), "unit": "utt/s",
        "batch_per_gpu": batch, "steps": steps, "final_loss": loss, "gpu_launches_per_step": launches / steps,
        "collective": (f"nccl all_reduce(SUM) of {(model.weight_count() + 1) * 4 / 1e6:.2f} MB flat fp32 gradients + found-inf "
                       f"flag per step, two buckets: {(model.weight_count() + 1 - trainer.bucket) * 4 / 1e6:.2f} MB (GRU + head, "
                       f"side stream, under the conv backward) then {trainer.bucket * 4 / 1e6:.2f} MB (conv); all_reduce_ms = what "
                       "the step still waits for after the conv backward" if ctx.world > 1 and trainer._graph is not None else
                       (f"nccl all_reduce(SUM) of {(model.weight_count() + 1) * 4 / 1e6:.2f} MB flat fp32 gradients + found-inf flag"
                        if ctx.world > 1 else "none at world size 1 (the all-reduce is skipped)")),
        "all_reduce_ms": round(coll, 4), "all_reduce_share": round(coll / (ms / steps), 4) if ms > 0 else None,
        "skipped_steps": trainer.skipped_steps,
        "cuda_graph": (f"{'three' if ctx.world > 1 else 'two'} graphs per step (split at the gradient all-reduces), {trainer._graph['launches']} kernels of "
                       "libsir_b200 per replay; step count, bias corrections, loss scale and dropout offset live on the device"
                       if trainer._graph is not None else "off (eager launches)"),
        "e2e": {"value": batch * ctx.world * steps / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": batch * 64 * OUT_FRAMES * 4 + batch * 8,
                "d2h_bytes_per_step": 8, "api": "DataParallelTrainer.step(features, labels) from pinned host buffers, loss read back"},
        "stages_ms": {k: round(v[0] / k_prof, 4) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])[:12]},
    }
    if cpu_ref and ctx.rank == 0 and ctx.world == 1:
        os.sched_setaffinity(0, ctx.all_cpus)
        arm = CpuTrainArm(ctx.synth, batch)
        arm.step()
        t0 = time.perf_counter()
        k = 0
        while k < 3 or (time.perf_counter() - t0 < 6.0 and k < 50):
            arm.step()
            k += 1
        s = (time.perf_counter() - t0) / k
        out["cpu_reference_step"] = {"ms_per_step": s * 1e3, "value": batch / s, "unit": "utt/s", "cores": arm.threads,
                                     "kind": "port", "sample": f"{k} fp32 training steps of batch {batch} on the host cores "
                                                               "(oracle/torch_port.ClassifierPort + torch.optim.Adam)"}
        if ctx.bound_cpus:
            os.sched_setaffinity(0, ctx.bound_cpus)
    return out


def run_training(ctx, out):
    args, wl = ctx.args, WORKLOADS["config4"]
    ctx.sampler.start()
    tr = measure_training(ctx, args.steps, args.warmup, batch=args.batch or wl["batch"], cpu_ref=not args.no_cpu_baseline)
    clocks = ctx.sampler.stop()
    if ctx.rank == 0:
        flops = 3 * 400.6e6 * tr["batch_per_gpu"]                       # forward + ~2x backward
        ach = flops / (tr["ms_per_step"] * 1e-3) / 1e12
        line = {
            "metric": "utterances/sec (training step)", "value": tr["value"], "unit": "utt/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tr["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": tr["batch_per_gpu"],
                       "parallelism": f"data parallel x{ctx.world}, one flat gradient all-reduce per step",
                       "l2_policy": "the step's working set (weights 13 MB, activations) is L2-resident by nature; batches "
                                    "are gathered from a 210 MB HBM-resident feature cache", "host_binding": ctx.host_binding()},
            "clocks": clocks, "gpu_launches": int(round(tr["gpu_launches_per_step"] * args.steps)),
            "e2e": tr["e2e"], "train": tr,
            "roofline": {"bound": "tensor", "achieved": round(ach, 3), "peak": ctx.peaks["tflops"], "unit": "TFLOP/s",
                         "frac": round(ach / ctx.peaks["tflops"], 5), "traffic": None, "kernel": "whole training step",
                         "note": "batch 16 is launch / latency bound (dozens of dependent launches per step); algorithmic "
                                 "FLOPs = 3 x 400.6 MFLOP per utterance"},
        }
        if "cpu_reference_step" in tr:
            c = tr["cpu_reference_step"]
            line["cpu_baseline"] = {"value": c["value"], "unit": "utt/s", "cores": c["cores"], "kind": c["kind"],
                                    "sample": c["sample"]}
        out.emit(line)


# ------------------------------------------------------------------------------------------------------------------
# feature precompute (config3)
# ------------------------------------------------------------------------------------------------------------------
def run_precompute(ctx, out):
    args, wl = ctx.args, WORKLOADS["config3"]
    train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")
    a, b = train.shard_range(args.batch or wl["total"], ctx.rank, ctx.world)
    n, L, n_mels = b - a, wl["samples"], wl["n_mels"]
    T = 1 + L // 512
    extractor = ctx.pre.AudioFeatureExtractor()
    base = torch.from_numpy(synth_batch(ctx.synth, 1000 + ctx.rank, 256, L))
    host_pcm = to_pcm16(base).repeat((n + 255) // 256, 1)[:n].contiguous().pin_memory()     # the shard as a WAV corpus holds it
    gains = torch.linspace(0.5, 1.0, n, device="cuda")[:, None]
    waves = (base.cuda().repeat((n + 255) // 256, 1)[:n] * gains).contiguous()               # fp32, HBM-resident shard
    feats = torch.empty((n, n_mels, T), device="cuda")

    def run_steps(k):
        for _ in range(k):
            extractor.extract_batch(waves, max_duration=5.0, out=feats)

    run_steps(args.warmup)
    ctx.sampler.start()
    l0 = ctx.native.launch_count()
    ms, _ = ctx.timed(lambda: run_steps(args.steps))
    launches = ctx.native.launch_count() - l0

    # e2e: int16 PCM chunks from pinned host memory -> features back in pinned host memory (what the cache writer keeps)
    chunk = 2048
    n_chunks = (n + chunk - 1) // chunk
    d_in = [torch.empty((chunk, L), dtype=torch.int16, device="cuda") for _ in range(2)]
    d_out = [torch.empty((chunk, n_mels, T), device="cuda") for _ in range(2)]
    host_feats = torch.empty((n, n_mels, T), dtype=torch.float32).pin_memory()
    streams = [torch.cuda.Stream() for _ in range(2)]

    def e2e_pass():
        t0 = time.perf_counter()
        for c in range(n_chunks):
            s, k = streams[c % 2], c % 2
            lo, hi = c * chunk, min(n, (c + 1) * chunk)
            with torch.cuda.stream(s):
                d_in[k][:hi - lo].copy_(host_pcm[lo:hi], non_blocking=True)
                extractor.extract_batch(d_in[k][:hi - lo], max_duration=5.0, out=d_out[k][:hi - lo])
                host_feats[lo:hi].copy_(d_out[k][:hi - lo], non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def e2e_run():
        ctx.barrier()
        ctx.sampler.mark()
        t = sum(e2e_pass() for _ in range(max(1, args.steps // 2)))
        ctx.barrier()
        ctx.sampler.mark()
        return t

    e2e_s, e2e_times, warm_used = stable_repeats(e2e_run, min(args.e2e_repeats, 3), agree=ctx.reduce_max)
    e2e_steps = max(1, args.steps // 2)
    clocks = ctx.sampler.stop()
    check = float((host_feats[:64].cuda() - extractor.extract_batch(host_pcm[:64].cuda(), max_duration=5.0)).abs().max())

    ctx.native.profile_enable(True)
    run_steps(args.steps)
    stages = ctx.native.profile_read()
    ctx.native.profile_enable(False)
    ms, e2e_s = ctx.reduce_max(ms), ctx.reduce_max(e2e_s)
    total = int(ctx.reduce_sum(n))
    if ctx.rank == 0:
        bytes_per_utt = 4 * L + 4 * n_mels * T
        stage_out, rooflines = stage_rooflines(ctx, stages, args.steps, n, n_mels, bytes_per_utt, False)
        fr = next((r for r in rooflines if r["kernel"] == "logmel_frontend_kernel"), None)
        line = {
            "metric": "utterances/sec (features)", "value": total * args.steps / (ms * 1e-3), "unit": "utt/s",
            "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "utterances_total": total, "utterances_this_rank": n, "samples": L,
                       "parallelism": f"contiguous shards x{ctx.world}, no collective",
                       "l2_policy": f"one pass reads {n * L * 4 / 1e6:.0f} MB per rank (> 126 MB L2)",
                       "host_binding": ctx.host_binding()},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total * e2e_steps / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": n * L * 2,
                    "d2h_bytes_per_step": n * n_mels * T * 4,
                    "api": f"AudioFeatureExtractor.extract_batch over {n_chunks} chunks of {chunk} int16 PCM utterances from pinned "
                           f"host memory, features copied back to pinned host memory, two streams; max |diff| vs the "
                           f"device-resident call {check:.1e}",
                    "seconds_of_each_repeat": [round(t, 4) for t in e2e_times], "warm_runs": warm_used},
            "roofline": fr, "frontend_roofline": fr, "rooflines": rooflines, "stages": stage_out,
        }
        if not args.no_cpu_baseline and ctx.world == 1:
            os.sched_setaffinity(0, ctx.all_cpus)
            arm = CpuArm(ctx.synth, 256, L, n_mels, classifier=False)
            arm.step()
            k, fs, ws = arm.run_for(args.cpu_seconds)
            line["cpu_baseline"] = {"value": arm.n * k / (fs + ws), "unit": "utt/s", "cores": arm.threads, "kind": arm.kind,
                                    "sample": arm.describe(k, fs, ws) + "; 256-utterance sample of the 30,043, rate scales linearly",
                                    "host_cpus": os.cpu_count() or 1}
        elif ctx.world > 1:
            line["cpu_baseline"] = None
            line["cpu_baseline_note"] = "timed on rank 0 at N = 1 only (see the N = 1 line)"
        out.emit(line)


# ------------------------------------------------------------------------------------------------------------------
# inference (config2, config5)
# ------------------------------------------------------------------------------------------------------------------
def run_inference(ctx, out):
    args = ctx.args
    wl = WORKLOADS[args.workload]
    native, rank, world = ctx.native, ctx.rank, ctx.world
    B, L, n_mels, max_duration = args.batch or wl["batch"], wl["samples"], wl["n_mels"], wl["max_duration"]
    extractor = ctx.pre.AudioFeatureExtractor(n_mels=n_mels)                       # the reference-facing objects (public API)
    sd = ctx.synth.make_weights(1234, NUM_CLASSES, n_mels)
    model = ctx.models.CNNAudioGRU(NUM_CLASSES, n_mels=n_mels)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    model = model.cuda().eval()

    def step_device(wave, feats):
        """One pass of the hot path with inputs already in HBM."""
        extractor.extract_batch(wave, max_duration=max_duration, out_frames=OUT_FRAMES, out=feats)
        return model(feats)

    # rotating input batches whose total size exceeds the 126 MB L2
    batch_mb = B * L * 4 / 1e6
    n_rot = max(2, int(np.ceil(190.0 / batch_mb)))
    host = torch.from_numpy(synth_batch(ctx.synth, 1000 + rank, B, L)).pin_memory()
    dev_waves = [(host.cuda() * (1.0 - 0.1 * (i % 5))).contiguous() for i in range(n_rot)]
    feats = torch.empty((B, n_mels, OUT_FRAMES), device="cuda")
    stream = ctx.stream
    # consecutive steps (independent batches) alternate over `--streams` CUDA streams, each with its own feature
    # buffer and - inside the model handle - its own workspace: the latency-bound GRU recurrence of step i overlaps
    # the frontend and conv stack of step i+1.  Every step still runs every kernel; all of them finish inside the
    # timed region (the timing stream waits for every worker stream before the closing event).
    workers = [torch.cuda.Stream() for _ in range(max(1, args.streams))]
    feats_w = [torch.empty((B, n_mels, OUT_FRAMES), device="cuda") for _ in workers]

    def run_steps(n):
        for w in workers:
            w.wait_stream(stream)
        res = None
        for i in range(n):
            k = i % len(workers)
            with torch.cuda.stream(workers[k]):
                res = step_device(dev_waves[i % n_rot], feats_w[k])
        for w in workers:
            stream.wait_stream(w)
        return res

    for i in range(args.warmup):
        step_device(dev_waves[i % n_rot], feats)
    run_steps(max(args.warmup, 2 * len(workers)))
    ctx.barrier()

    # ---- value: K steps, inputs resident in HBM ------------------------------------------------------------
    ctx.sampler.start()
    launches0 = native.launch_count()
    dev_ms, _ = ctx.timed(lambda: run_steps(args.steps))
    launches = native.launch_count() - launches0

    # ---- e2e: host buffers, H2D + pipeline + D2H each step, public API -------------------------------------
    # IntentPipeline.submit/collect: pinned host waveforms in, pinned host logits out; up to `depth` batches in flight on
    # their own streams, so the H2D copy of batch k+1 overlaps the compute of batch k.
    pipe = ctx.pipe_mod.IntentPipeline(extractor, model, sub_batches=args.sub_batches, out_frames=OUT_FRAMES,
                                       max_duration=max_duration, depth=args.depth)
    host_pcm = to_pcm16(host).pin_memory()                 # what a 16-bit WAV holds / torchaudio.load starts from
    pipe.reserve(B, L)                                     # slot buffers for both dtypes, before any timed loop

    def e2e_loop(n, src):
        """n steps, each with its own H2D copy and D2H read; up to `depth` batches in flight, all drained before return."""
        pending, res = [], None
        t_submit = t_wait = 0.0
        for _ in range(n):
            t0 = time.perf_counter()
            pending.append(pipe.submit(src))
            t1 = time.perf_counter()
            if len(pending) == args.depth:
                res = pipe.collect(pending.pop(0))
            t_submit += t1 - t0
            t_wait += time.perf_counter() - t1
        while pending:
            res = pipe.collect(pending.pop(0))
        last["host_ms"] = (round(t_submit * 1e3 / n, 4), round(t_wait * 1e3 / n, 4))
        return res

    last = {}

    def e2e_run(src, mark):
        ctx.barrier()
        if mark:
            ctx.sampler.mark()
        t0 = time.perf_counter()
        last["logits"] = e2e_loop(args.steps, src)
        ctx.barrier()
        t = time.perf_counter() - t0
        if mark:
            ctx.sampler.mark()
        return t

    e2e_s, e2e_times, e2e_warm = stable_repeats(lambda: e2e_run(host_pcm, True), args.e2e_repeats, agree=ctx.reduce_max)
    pcm_logits = last["logits"].clone()
    e2e_host_ms = last["host_ms"]
    clocks = ctx.sampler.stop()
    e2e_f32_s, e2e_f32_times, e2e_f32_warm = stable_repeats(lambda: e2e_run(host, False), args.e2e_repeats, agree=ctx.reduce_max)
    # the same kernels on the same bytes give the same result as the device-resident path
    pcm_dev = (host_pcm.cuda().to(torch.float32) / 32768.0).contiguous()
    e2e_check = float((pcm_logits.cuda() - step_device(pcm_dev, feats)).abs().max())
    h2d_pcm_ms = h2d_rate(ctx, host_pcm)
    h2d_f32_ms = h2d_rate(ctx, host)

    # ---- per-stage device times: separate pass with events around every stage ------------------------------
    # three passes of K steps; per stage the MEDIAN of the three K-step totals (one pass on a power-capped box once timed a
    # single stage 55 % above every other run)
    native.profile_enable(True)
    passes = []
    for _ in range(3):
        for i in range(args.steps):
            step_device(dev_waves[i % n_rot], feats)
        passes.append(native.profile_read())
    native.profile_enable(False)
    stages = {k: (sorted(p[k][0] for p in passes)[1], passes[0][k][1]) for k in passes[0]}

    dev_ms, e2e_s, e2e_f32_s = ctx.reduce_max(dev_ms), ctx.reduce_max(e2e_s), ctx.reduce_max(e2e_f32_s)
    train_obj = None
    if args.workload == "config2" and not args.no_train:
        del dev_waves, feats_w
        train_obj = measure_training(ctx, args.train_steps, args.warmup, cpu_ref=not args.no_cpu_baseline)
    if rank == 0:
        total_utts = B * world * args.steps
        fe_bytes = 4 * min(L, int(max_duration * 16000) if max_duration else L) + 4 * n_mels * OUT_FRAMES
        stage_out, rooflines = stage_rooflines(ctx, stages, args.steps, B, n_mels, fe_bytes,
                                               args.workload == "config2" and B == 256)
        slowest = max(rooflines, key=lambda r: r["ms"]) if rooflines else None
        fr = next((r for r in rooflines if r["kernel"] == "logmel_frontend_kernel"), None)
        # `roofline` is the frontend's: BASELINE.json's metric names it ("frontend HBM GB/s vs peak") and it is the one
        # HBM-bound kernel of the step.  The slowest single stage is reported beside it, every stage in `rooflines`.
        roofline = dict(fr, choice="the kernel BASELINE.json's metric names (frontend HBM GB/s vs peak); the slowest "
                                   f"single stage is {slowest['kernel']} ({slowest['ms']} ms), see slowest_stage_roofline "
                                   "and rooflines") if fr else slowest

        def copy_bound(ms, nbytes):
            return {"h2d_gbs_per_gpu": round(nbytes / (ms * 1e-3) / 1e9, 2), "utt_s": round(B * world / (ms * 1e-3)),
                    "note": "pinned host -> device copy of one batch per step, all ranks copying at once: the ceiling of this "
                            "end-to-end number on this host"}

        line = {
            "metric": "utterances/sec (features+forward)", "value": total_utts / (dev_ms * 1e-3), "unit": "utt/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "samples": L,
                       "parallelism": f"batch-sharded x{world}, no collective",
                       "concurrency": f"consecutive steps alternate over {len(workers)} CUDA streams (value) / {args.depth} "
                                      "pipeline slots with their own streams (e2e); `stages` are timed serially on one "
                                      "stream (median of three K-step passes), so their sum exceeds ms_per_step",
                       "host_binding": ctx.host_binding(),
                       "l2_policy": f"{n_rot} rotating input batches ({n_rot * batch_mb:.0f} MB > 126 MB L2)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total_utts / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": B * L * 2,
                    "d2h_bytes_per_step": B * NUM_CLASSES * 4,
                    "api": f"IntentPipeline.submit/collect: pinned host int16 PCM waveforms (scaled by 1/32768 on the device, "
                           f"sir_frontend_forward_pcm16) -> pinned host logits, depth {args.depth} in flight, "
                           f"{args.sub_batches} sub-batch(es) per batch; max |logit diff| vs the device-resident path {e2e_check:.1e}",
                    "repeats": args.e2e_repeats, "statistic": "median of the repeats, each K steps, after warm runs until two "
                                                              "consecutive runs agree within 3 %",
                    "warm_runs": e2e_warm, "ms_per_step_of_each_repeat": [round(t * 1e3 / args.steps, 4) for t in e2e_times],
                    "fraction_of_value": round((total_utts / e2e_s) / (total_utts / (dev_ms * 1e-3)), 4),
                    "host_ms_per_step": {"enqueue": e2e_host_ms[0], "waiting_in_collect": e2e_host_ms[1],
                                         "note": "rank 0, last repeat: CPU time inside submit() (enqueueing copies and "
                                                 "launches) and time blocked in collect(); enqueue close to the step time "
                                                 "means the host thread, not the GPU or the copy, bounds this number"},
                    "h2d_copy_bound": copy_bound(h2d_pcm_ms, B * L * 2)},
            "e2e_fp32": {"value": total_utts / e2e_f32_s, "unit": "utt/s", "h2d_bytes_per_step": B * L * 4,
                         "d2h_bytes_per_step": B * NUM_CLASSES * 4, "api": "the same submit/collect loop with fp32 host buffers",
                         "warm_runs": e2e_f32_warm,
                         "ms_per_step_of_each_repeat": [round(t * 1e3 / args.steps, 4) for t in e2e_f32_times],
                         "h2d_copy_bound": copy_bound(h2d_f32_ms, B * L * 4)},
            "roofline": roofline, "frontend_roofline": fr, "slowest_stage_roofline": slowest, "rooflines": rooflines,
            "stages": stage_out,
        }
        if train_obj is not None:
            line["train"] = train_obj
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, ctx.all_cpus)                     # the CPU arm gets every host core again
            n_cpu = B if args.workload == "config2" else 16
            arm = CpuArm(ctx.synth, n_cpu, L, n_mels, max_duration=max_duration)
            arm.step()
            n_steps, fs, ws = arm.run_for(args.cpu_seconds)
            line["cpu_baseline"] = {"value": arm.n * n_steps / (fs + ws), "unit": "utt/s", "cores": arm.threads,
                                    "kind": arm.kind, "sample": arm.describe(n_steps, fs, ws),
                                    "host_cpus": os.cpu_count() or 1}
            # the same path with ONE thread (the reference pipeline exports OMP_NUM_THREADS=1,
            # run_pipeline.py:42) and the CPU's best case (the whole batch in one mel_transform call instead of the
            # per-utterance loop), so that the ratio is not read as Python overhead; a few seconds each
            if args.workload == "config2" and arm.ref_extractor is not None:
                try:
                    with torch.no_grad():
                        w = arm.pcm.to(torch.float32) / 32768.0
                        ex = arm.ref_extractor

                        def batched():
                            m = ex.amplitude_to_db(ex.mel_transform(w))
                            m = (m - m.mean(dim=(1, 2), keepdim=True)) / (m.std(dim=(1, 2), keepdim=True) + 1e-5)
                            arm.model(torch.nn.functional.pad(m, (0, OUT_FRAMES - m.shape[2])))
                        batched()
                        t0 = time.perf_counter()
                        k = 0
                        while k < 2 or (time.perf_counter() - t0 < 2.0 and k < 20):
                            batched()
                            k += 1
                        line["cpu_baseline"]["batched_features_value"] = arm.n * k / (time.perf_counter() - t0)
                        torch.set_num_threads(1)
                        sub = CpuArm(ctx.synth, 32, L, n_mels, max_duration=max_duration, threads=1)
                        sub.step()
                        k1, f1, w1 = sub.run_for(3.0, min_steps=1, max_steps=8)
                        line["cpu_baseline"]["one_thread_value"] = sub.n * k1 / (f1 + w1)
                        line["cpu_baseline"]["variants"] = ("batched_features_value: one mel_transform call for the whole batch + "
                                                            "batched forward, all cores; one_thread_value: the per-utterance "
                                                            "loop + forward on 32 utterances with torch.set_num_threads(1)")
                except Exception as e:  # noqa: BLE001 - the variants are optional context
                    line["cpu_baseline"]["variants"] = f"not measured: {type(e).__name__}: {e}"
                finally:
                    torch.set_num_threads(os.cpu_count() or 1)
        elif world > 1:
            line["cpu_baseline"] = None
            line["cpu_baseline_note"] = "timed on rank 0 at N = 1 only (see the N = 1 line): no CPU work while other ranks hold GPUs"
        out.emit(line)


def main():
    out = JsonStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU per step (config3: total utterances)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="config2: skip the `train` object")
    ap.add_argument("--no-graph", action="store_true", help="training step: eager launches instead of CUDA-graph replays")
    ap.add_argument("--train-steps", type=int, default=50, help="config2: steps of the `train` object")
    ap.add_argument("--sub-batches", type=int, default=1, help="sub-batches of the host-buffer pipeline (e2e)")
    ap.add_argument("--depth", type=int, default=4, help="batches in flight in the host-buffer pipeline (e2e)")
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams consecutive steps alternate over (value)")
    ap.add_argument("--e2e-repeats", type=int, default=5, help="runs of K end-to-end steps; the median is reported")
    ap.add_argument("--numa-bind", type=int, default=1, help="bind each rank to its GPU's CPU cores before allocating pinned buffers")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work spent on the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps is None:
        args.steps = WORKLOADS[args.workload]["steps"]

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), out)
        return
    ctx = Ctx(args)
    {"config2": run_inference, "config5": run_inference, "config3": run_precompute, "config4": run_training}[args.workload](ctx, out)
    ctx.finish()


if __name__ == "__main__":
    main()
