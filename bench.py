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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (hbm copy; bf16 sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback of B200_PROFILING.md"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batch(native_synth, seed, batch):
    """Speech-like rows are expensive to synthesise on the host; tile 32 distinct utterances with per-row gains."""
    base = native_synth.speech_like(seed, 32, SAMPLES)
    reps = (batch + 31) // 32
    gains = np.linspace(0.25, 1.0, reps * 32, dtype=np.float32)[:, None]
    return (np.tile(base, (reps, 1)) * gains)[:batch]


def cpu_reference_arm(native_synth, n_utts, threads=None):
    """The reference's CPU path on `n_utts` utterances -> (utt/s, seconds, description)."""
    from oracle.torch_port import ClassifierPort, FeaturePort, load_numpy_state
    if threads:
        torch.set_num_threads(threads)
    waves = torch.from_numpy(synth_batch(native_synth, 99, n_utts))
    fp = FeaturePort()
    model = load_numpy_state(ClassifierPort(NUM_CLASSES).eval(), native_synth.make_weights(1234))
    t0 = time.perf_counter()
    feats = fp.batch_padded(waves, target=OUT_FRAMES)          # one call per utterance, like the reference loop
    t1 = time.perf_counter()
    with torch.no_grad():
        logits = model(feats)
    t2 = time.perf_counter()
    return n_utts / (t2 - t0), (t1 - t0, t2 - t1), int(logits.argmax(1)[0])


def run_reference(args, rank, world):
    if rank != 0:
        return
    native_synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    per_step = 64
    for _ in range(args.warmup):
        cpu_reference_arm(native_synth, 16)
    t0 = time.perf_counter()
    feat_s = fwd_s = 0.0
    for _ in range(args.steps):
        _, (a, b), _ = cpu_reference_arm(native_synth, per_step)
        feat_s += a
        fwd_s += b
    total = feat_s + fwd_s
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": "utterances/sec (features+forward)", "value": value, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 256 utt x 3 s @16 kHz, config.yaml model (64 mel, 200 frames, 31 classes)",
                   "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "utt/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{per_step} utterances x 3 s per step: per-utterance torchaudio MelSpectrogram+"
                                   f"AmplitudeToDB+normalise loop ({feat_s / args.steps:.2f} s) + CNNAudioGRU fp32 "
                                   f"forward in one batch ({fwd_s / args.steps:.2f} s); wall {time.perf_counter() - t0:.1f} s",
                         "host_cpus": cores, "torch": torch.__version__},
        "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="utterances per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        logits = step_device(dev_waves[i % n_rot], feats)
    barrier()

    # ---- value: K steps, inputs resident in HBM ------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        logits = step_device(dev_waves[i % n_rot], feats)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = native.launch_count() - launches0

    # ---- e2e: host buffers, H2D + pipeline + D2H each step, public API -------------------------------------
    host_logits = torch.empty((B, NUM_CLASSES), dtype=torch.float32).pin_memory()
    dev_in = torch.empty((B, SAMPLES), device="cuda")
    for _ in range(3):
        dev_in.copy_(host, non_blocking=True)
        host_logits.copy_(step_device(dev_in, feats), non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        dev_in.copy_(host, non_blocking=True)
        host_logits.copy_(step_device(dev_in, feats), non_blocking=True)
        torch.cuda.current_stream().synchronize()              # the caller reads the step's result
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

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
        model_stages = {k: v for k, v in stage_out.items() if k in FLOPS_PER_UTT}
        dom = max(model_stages, key=lambda k: model_stages[k]["ms_per_step"]) if model_stages else None
        roofline = None
        if dom:
            ach = model_stages[dom]["tflops"]
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": round(ach / peaks["tflops"], 5), "traffic": None, "peak_source": peaks["source"],
                        "share_of_step": round(model_stages[dom]["ms_per_step"] /
                                               sum(v["ms_per_step"] for v in stage_out.values()), 4)}
        fr = None
        if "logmel_frontend_kernel" in stage_out:
            ms = stage_out["logmel_frontend_kernel"]["ms_per_step"]
            gbs = FRONTEND_BYTES_PER_UTT * B / (ms * 1e-3) / 1e9
            fr = {"kernel": "logmel_frontend_kernel", "bound": "hbm", "achieved": round(gbs, 2), "peak": peaks["hbm_gbs"],
                  "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 5), "traffic": None,
                  "bytes_per_utt": FRONTEND_BYTES_PER_UTT, "peak_source": peaks["source"]}
        line = {
            "metric": "utterances/sec (features+forward)", "value": value, "unit": "utt/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 256 utt x 3 s @16 kHz per GPU, config.yaml model "
                                   "(64 mel, 200 frames, 31 classes), seeded random-init weights",
                       "batch_per_gpu": B, "samples": SAMPLES, "parallelism": f"batch-sharded x{world}, no collective",
                       "l2_policy": f"{n_rot} rotating input batches ({n_rot * B * SAMPLES * 4 / 1e6:.0f} MB > 126 MB L2)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total_utts / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": B * SAMPLES * 4,
                    "d2h_bytes_per_step": B * NUM_CLASSES * 4,
                    "api": "AudioFeatureExtractor.extract_batch + CNNAudioGRU.forward on pinned host buffers"},
            "roofline": roofline, "frontend_roofline": fr, "stages": stage_out,
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cpu_reference_arm(native_synth, 8)
            v, (fs, ws), _ = cpu_reference_arm(native_synth, 64)
            line["cpu_baseline"] = {"value": v, "unit": "utt/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"64 of the 256 utterances: per-utterance torchaudio feature loop "
                                              f"{fs:.2f} s + CNNAudioGRU fp32 batched forward {ws:.2f} s",
                                    "host_cpus": cores}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
