#!/usr/bin/env python
"""bench.py - utterances/sec (features + forward) of the B200-native hot path, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): a synthetic FSC-shaped inference batch of 256 utterances x 3 s @16 kHz per
GPU, config.yaml model (64 mels, 200 frames, 31 classes), random-init weights of that architecture.  A step is
one pass of the hot path over one batch: fused log-mel frontend + CNNAudioGRU forward -> logits.  Utterances are
independent, so N GPUs run N shards with no data-path collective ("scaling": "weak").

 value  : whole-job utterances/s with inputs resident in HBM, K back-to-back steps between CUDA events, max over
          ranks.  Steps rotate over input buffers whose total size exceeds the 126 MB L2.
 e2e    : the same metric through the public Python API with HOST (pinned) buffers: H2D copy of the waveforms,
          the pipeline, D2H read of the logits, every step, wall clock around the synchronised region.
 roofline / frontend_roofline / stages : per-kernel device times from a separate K-step pass with CUDA events
          around each stage on the launching stream (sir_profile_*), algorithmic flops/bytes per DESIGN.md.
 cpu_baseline : the reference's CPU path (oracle/torch_port.py: the same torchaudio / torch.nn calls the
          reference makes, per-utterance feature loop + batched fp32 forward) on this box's host cores, rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 256
SAMPLES = 48000
N_MELS, OUT_FRAMES, NUM_CLASSES = 64, 200, 31
FRAMES = 1 + SAMPLES // 512
FLOPS_PER_UTT = {  # SURVEY.md 8 a11 (2 flops per MAC)
    "conv1_bn_relu_pool": 7.37e6, "conv2_bn_relu_pool": 117.96e6, "conv3_bn_relu_pool": 117.96e6,
    "gru_l0_input_gemm": 78.64e6, "gru_l0_recurrence": 19.66e6, "gru_l1_input_gemm": 39.32e6,
    "gru_l1_recurrence": 19.66e6, "attention_fc": 0.06e6,
}
FRONTEND_BYTES_PER_UTT = 4 * SAMPLES + 4 * N_MELS * OUT_FRAMES      # reads the waveform once, writes [64,200] once
# conv2 / conv3 / the GRU projections run as 3-pass fp16 hi/lo splits: the tensor pipe executes 3x these FLOPs.
SPLIT_PASSES = {"conv2_bn_relu_pool": 3, "conv3_bn_relu_pool": 3, "gru_l0_input_gemm": 3, "gru_l1_input_gemm": 3,
                "gru_l0_recurrence": 3, "gru_l1_recurrence": 3}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of THIS workload
# (256 utterances; profiles/r1_summary.md names the capture of each row).  None: not captured for this build.
NCU_TRAFFIC_B256 = {   # captures r1g (profiles/r1g_ncu_*.txt)
    "logmel_frontend_kernel": 54.17e6 + 0.85e6, "conv1_bn_relu_pool": 22.68e6 + 63.17e6,
    "conv2_bn_relu_pool": 104.97e6 + 30.97e6, "conv3_bn_relu_pool": 52.76e6 + 4.88e6,
    "gru_l0_input_gemm": 32.53e6 + 3.28e6, "gru_l1_input_gemm": 16.28e6 + 0.34e6,
    "gru_l0_recurrence": 40.92e6 + 0.27e6, "gru_l1_recurrence": 40.92e6 + 0.03e6,     # r1j capture: unchanged
}


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


def synth_batch(native_synth, seed, batch):
    """Speech-like rows are expensive to synthesise on the host; tile 32 distinct utterances with per-row gains."""
    base = native_synth.speech_like(seed, 32, SAMPLES)
    reps = (batch + 31) // 32
    gains = np.linspace(0.25, 1.0, reps * 32, dtype=np.float32)[:, None]
    return (np.tile(base, (reps, 1)) * gains)[:batch]


class CpuArm:
    """The reference's CPU path (oracle/torch_port.py: the torchaudio / torch.nn calls the reference makes) on the
    host cores: per-utterance feature loop like scripts/precompute_features.py:124-130, then CNNAudioGRU.eval()
    fp32 forward in one batch.  Test/bench infrastructure only - never on the product path."""

    def __init__(self, native_synth, n_utts, threads=None):
        from oracle.torch_port import ClassifierPort, FeaturePort, load_numpy_state
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.n = n_utts
        self.waves = torch.from_numpy(synth_batch(native_synth, 99, n_utts))
        self.fp = FeaturePort()
        self.model = load_numpy_state(ClassifierPort(NUM_CLASSES).eval(), native_synth.make_weights(1234))

    def step(self):
        """One pass over the n_utts batch -> (feature seconds, forward seconds)."""
        t0 = time.perf_counter()
        feats = self.fp.batch_padded(self.waves, target=OUT_FRAMES)   # one call per utterance, like the reference loop
        t1 = time.perf_counter()
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


def run_reference(args, rank, world, out):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, rank 0 only.
    One step = the full configs[1] batch (256 utterances x 3 s)."""
    if rank != 0:
        return
    native_synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
    arm = CpuArm(native_synth, BATCH_PER_GPU)
    for _ in range(min(args.warmup, 3)):
        arm.step()
    t0 = time.perf_counter()
    feat_s = fwd_s = 0.0
    for _ in range(args.steps):
        a, b = arm.step()
        feat_s += a
        fwd_s += b
    total = feat_s + fwd_s
    value = arm.n * args.steps / total
    line = {
        "impl": "reference", "metric": "utterances/sec (features+forward)", "value": value, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 256 utt x 3 s @16 kHz, config.yaml model (64 mel, 200 frames, 31 classes)",
                   "batch_per_step": arm.n},
        "cpu_baseline": {"value": value, "unit": "utt/s", "cores": arm.threads, "kind": "port",
                         "sample": f"{args.steps} steps x {arm.n} utterances x 3 s: per-utterance torchaudio "
                                   f"MelSpectrogram+AmplitudeToDB+normalise loop ({feat_s / args.steps:.3f} s/step) + "
                                   f"CNNAudioGRU fp32 forward in one batch ({fwd_s / args.steps:.3f} s/step); "
                                   f"wall {time.perf_counter() - t0:.1f} s",
                         "host_cpus": os.cpu_count() or 1, "torch": torch.__version__},
        "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    out.emit(line)


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


def main():
    out = JsonStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="utterances per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub-batches", type=int, default=1, help="sub-batches of the host-buffer pipeline (e2e)")
    ap.add_argument("--depth", type=int, default=4, help="batches in flight in the host-buffer pipeline (e2e)")
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams consecutive steps alternate over (value)")
    ap.add_argument("--e2e-repeats", type=int, default=5, help="runs of K end-to-end steps; the median is reported")
    ap.add_argument("--numa-bind", type=int, default=1, help="bind each rank to its GPU's CPU cores before allocating pinned buffers")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work spent on the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    native = importlib.import_module("speech-intent-recognizer_b200._native")
    native_synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
    pre = importlib.import_module("speech-intent-recognizer_b200.scripts.precompute_features")
    models = importlib.import_module("speech-intent-recognizer_b200.models.models")

    pipe_mod = importlib.import_module("speech-intent-recognizer_b200.pipeline")
    all_cpus = os.sched_getaffinity(0)
    bound_cpus = pipe_mod.bind_host_to_gpu(local) if args.numa_bind else None
    B = args.batch
    extractor = pre.AudioFeatureExtractor()                      # the reference-facing objects (public API)
    sd = native_synth.make_weights(1234)
    model = models.CNNAudioGRU(NUM_CLASSES)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    model = model.cuda().eval()

    def step_device(wave, feats):
        """One pass of the hot path with inputs already in HBM."""
        extractor.extract_batch(wave, max_duration=5.0, out_frames=OUT_FRAMES, out=feats)
        return model(feats)

    # synthetic inputs: 4 rotating batches of 49 MB -> 197 MB > L2 (126 MB)
    n_rot = 4
    host = torch.from_numpy(synth_batch(native_synth, 1000 + rank, B)).pin_memory()
    dev_waves = [(host.cuda() * (1.0 - 0.1 * i)).contiguous() for i in range(n_rot)]
    feats = torch.empty((B, N_MELS, OUT_FRAMES), device="cuda")
    stream = torch.cuda.current_stream()
    # consecutive steps (independent batches) alternate over `--streams` CUDA streams, each with its own feature
    # buffer and - inside the model handle - its own workspace: the latency-bound GRU recurrence of step i overlaps
    # the frontend and conv stack of step i+1.  Every step still runs every kernel; all of them finish inside the
    # timed region (the timing stream waits for every worker stream before the closing event).
    workers = [torch.cuda.Stream() for _ in range(max(1, args.streams))]
    feats_w = [torch.empty((B, N_MELS, OUT_FRAMES), device="cuda") for _ in workers]

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        logits = step_device(dev_waves[i % n_rot], feats)
    run_steps(max(args.warmup, 2 * len(workers)))
    barrier()

    # ---- value: K steps, inputs resident in HBM ------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record(stream)
    logits = run_steps(args.steps)
    e1.record(stream)
    barrier()
    sampler.mark()
    dev_ms = e0.elapsed_time(e1)
    launches = native.launch_count() - launches0

    # ---- e2e: host buffers, H2D + pipeline + D2H each step, public API -------------------------------------
    # IntentPipeline.infer_host: pinned host waveforms in, pinned host logits out; the H2D copy of sub-batch i+1
    # overlaps the frontend + conv stack of sub-batch i; it synchronises before returning (the caller reads logits).
    pipe = pipe_mod.IntentPipeline(extractor, model, sub_batches=args.sub_batches, out_frames=OUT_FRAMES, max_duration=5.0,
                                   depth=args.depth)

    def e2e_loop(n, src=None):
        """n steps, each with its own H2D copy and D2H read; up to `depth` batches in flight, all drained before return."""
        src = host if src is None else src
        pending, res = [], None
        for _ in range(n):
            pending.append(pipe.submit(src))
            if len(pending) == args.depth:
                res = pipe.collect(pending.pop(0))
        while pending:
            res = pipe.collect(pending.pop(0))
        return res

    def timed_e2e(src, mark):
        """Median wall-clock time of K end-to-end steps over `e2e_repeats` runs (host-side jitter moves single runs of a
        few tens of milliseconds by a lot when the loop is not copy-bound)."""
        e2e_loop(max(10, 2 * args.depth), src)
        times, res = [], None
        for _ in range(max(1, args.e2e_repeats)):
            barrier()
            if mark:
                sampler.mark()
            t0 = time.perf_counter()
            res = e2e_loop(args.steps, src)
            barrier()
            times.append(time.perf_counter() - t0)
            if mark:
                sampler.mark()
        e2e_runs.append([round(t * 1e3 / args.steps, 4) for t in times])
        return float(np.median(times)), res

    e2e_runs = []                                            # ms per step of every repeat: [fp32 runs, pcm16 runs]
    e2e_s, host_logits = timed_e2e(host, True)
    clocks = sampler.stop()
    e2e_check = float((host_logits.cuda() - step_device(dev_waves[0], feats)).abs().max())   # same kernels, same result
    # the same loop fed with 16-bit PCM host buffers (what a WAV file holds): half the PCIe bytes, scaled on the device
    host_pcm = (host * 32767.0).round().to(torch.int16).pin_memory()
    e2e_pcm_s, _ = timed_e2e(host_pcm, False)

    # the bound of the fp32 end-to-end number: pinned host -> device copy rate of one batch, all ranks copying at once
    d_probe = torch.empty_like(dev_waves[0])
    for _ in range(2):
        d_probe.copy_(host, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for _ in range(10):
        d_probe.copy_(host, non_blocking=True)
    c1.record(stream)
    barrier()
    h2d_ms = c0.elapsed_time(c1) / 10
    del d_probe

    # ---- per-stage device times: separate pass with events around every stage ------------------------------
    native.profile_enable(True)
    for i in range(args.steps):
        step_device(dev_waves[i % n_rot], feats)
    stages = native.profile_read()
    native.profile_enable(False)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_ms = reduce_max(dev_ms)
    e2e_s = reduce_max(e2e_s)
    e2e_pcm_s = reduce_max(e2e_pcm_s)
    h2d_ms = reduce_max(h2d_ms)
    if rank == 0:
        peaks = measured_peaks()
        total_utts = B * world * args.steps
        value = total_utts / (dev_ms * 1e-3)
        stage_out = {}
        for name, (ms, calls) in stages.items():
            per_step = ms / args.steps
            entry = {"ms_per_step": round(per_step, 5), "launch_groups_per_step": calls // args.steps}
            if name in FLOPS_PER_UTT:
                entry["tflops"] = round(FLOPS_PER_UTT[name] * B / (per_step * 1e-3) / 1e12, 3)
            stage_out[name] = entry
        step_ms = sum(v["ms_per_step"] for v in stage_out.values())

        def stage_roofline(name):
            """roofline object of one stage: algorithmic FLOPs or bytes / event-timed duration vs the measured peak."""
            v = stage_out[name]
            traffic = NCU_TRAFFIC_B256.get(name) if B == BATCH_PER_GPU else None
            base = {"kernel": name, "ms": v["ms_per_step"], "share_of_step": round(v["ms_per_step"] / step_ms, 4),
                    "traffic": traffic, "peak_source": peaks["source"]}
            if name == "logmel_frontend_kernel":
                gbs = FRONTEND_BYTES_PER_UTT * B / (v["ms_per_step"] * 1e-3) / 1e9
                return {**base, "bound": "hbm", "achieved": round(gbs, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": round(gbs / peaks["hbm_gbs"], 5), "bytes_per_utt": FRONTEND_BYTES_PER_UTT,
                        "note": "CUDA-core FFT: ~1,300 instructions per lane per frame cap the kernel near 30 % of the HBM roofline; latency bound below that (DESIGN.md section 4)"}
            if name == "conv1_bn_relu_pool":
                nbytes = (4 * N_MELS * OUT_FRAMES + 4 * 32 * (N_MELS // 2) * (OUT_FRAMES // 2)) * B
                gbs = nbytes / (v["ms_per_step"] * 1e-3) / 1e9
                return {**base, "bound": "hbm", "achieved": round(gbs, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": round(gbs / peaks["hbm_gbs"], 5)}
            ach = v.get("tflops", 0.0)
            out = {**base, "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                   "frac": round(ach / peaks["tflops"], 5)}
            if name in SPLIT_PASSES:
                out["tensor_pipe_flops_factor"] = SPLIT_PASSES[name]
                out["note"] = ("fp32-accurate 3-pass fp16 hi/lo split: the tensor pipe executes 3x the algorithmic FLOPs "
                               f"({round(3 * ach, 1)} TFLOP/s of fp16 MMA work)")
            if "recurrence" in name:
                out["note"] = "latency chain of 25 dependent time steps (one launch per layer, 3.6 us per step, two 32-utterance chains per cluster); " + out.get("note", "")
            return out

        rooflines = [stage_roofline(k) for k in stage_out]
        roofline = max(rooflines, key=lambda r: r["ms"]) if rooflines else None
        fr = next((r for r in rooflines if r["kernel"] == "logmel_frontend_kernel"), None)
        line = {
            "metric": "utterances/sec (features+forward)", "value": value, "unit": "utt/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 256 utt x 3 s @16 kHz per GPU, config.yaml model "
                                   "(64 mel, 200 frames, 31 classes), seeded random-init weights",
                       "batch_per_gpu": B, "samples": SAMPLES, "parallelism": f"batch-sharded x{world}, no collective",
                       "concurrency": f"consecutive steps alternate over {len(workers)} CUDA streams (value) / {args.depth} "
                                      "pipeline slots with their own streams (e2e); `stages` are timed serially on one "
                                      "stream, so their sum exceeds ms_per_step",
                       "host_binding": (f"rank pinned to the {len(bound_cpus)} CPU cores NVML lists for its GPU before "
                                        "allocating pinned buffers" if bound_cpus else "none"),
                       "l2_policy": f"{n_rot} rotating input batches ({n_rot * B * SAMPLES * 4 / 1e6:.0f} MB > 126 MB L2)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total_utts / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": B * SAMPLES * 4,
                    "d2h_bytes_per_step": B * NUM_CLASSES * 4,
                    "api": f"IntentPipeline.submit/collect (pinned host waveforms -> pinned host logits), depth {args.depth} "
                           f"in flight, {args.sub_batches} sub-batches per batch: H2D overlapped with frontend + conv "
                           f"stack and with the previous batch's GRU/head; max |logit diff| vs the device-resident "
                           f"path {e2e_check:.1e}",
                    "repeats": args.e2e_repeats, "statistic": "median of the repeats, each K steps",
                    "ms_per_step_of_each_repeat": e2e_runs[0],
                    "h2d_copy_bound": {"h2d_gbs_per_gpu": round(B * SAMPLES * 4 / (h2d_ms * 1e-3) / 1e9, 2),
                                       "utt_s": round(total_utts / args.steps / (h2d_ms * 1e-3)),
                                       "note": "pinned host -> device copy of one fp32 batch per step, all ranks copying "
                                               "at once: the ceiling of e2e.value on this host"}},
            "e2e_pcm16": {"value": total_utts / e2e_pcm_s, "unit": "utt/s", "h2d_bytes_per_step": B * SAMPLES * 2,
                          "d2h_bytes_per_step": B * NUM_CLASSES * 4,
                          "api": "the same submit/collect loop with int16 PCM host buffers (sir_frontend_forward_pcm16)",
                          "ms_per_step_of_each_repeat": e2e_runs[1]},
            "roofline": roofline, "frontend_roofline": fr, "rooflines": rooflines, "stages": stage_out,
        }
        if not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)                     # the CPU arm gets every host core again
            arm = CpuArm(native_synth, B)
            arm.step()
            n_steps, fs, ws = arm.run_for(args.cpu_seconds)
            line["cpu_baseline"] = {"value": arm.n * n_steps / (fs + ws), "unit": "utt/s", "cores": arm.threads,
                                    "kind": "port",
                                    "sample": f"{n_steps} passes over the {arm.n}-utterance batch ({fs + ws:.1f} s of CPU "
                                              f"work): per-utterance torchaudio feature loop {fs / n_steps:.3f} s + "
                                              f"CNNAudioGRU fp32 batched forward {ws / n_steps:.3f} s per pass",
                                    "host_cpus": os.cpu_count() or 1}
        out.emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
