"""Time the fused training step (DataParallelTrainer.step) on one GPU or under torchrun: ms/step and utt/s."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import synth, train_inputs
models = importlib.import_module("speech-intent-recognizer_b200.models.models")
train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")
import torch.distributed as dist

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sd = synth.make_weights(1234)
    m = models.CNNAudioGRU(31)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    m = m.cuda()
    tr = train.DataParallelTrainer(m, lr=5e-5, weight_decay=1e-4)
    x, y = train_inputs(seed=5, batch=B)
    x, y = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for _ in range(5):
        tr.step(x, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = tr.step(x, y)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"train step B={B}/gpu x{world}: {dt * 1e3:.3f} ms/step -> {B * world / dt:.0f} utt/s, loss {loss:.4f}, adam steps {tr.adam_steps}, skipped {tr.skipped_steps}")
    if world > 1:
        dist.destroy_process_group()

main()
