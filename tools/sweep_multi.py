"""BASELINE config 4: global batch sweep sharded over the GPUs of one box, per-frame logits all-gathered.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/sweep_multi.py --batches 8,64,256,1024

Every rank runs the encoder (CUDA-graph replay) on its contiguous slice of the global batch
(`distributed.partition`, shards padded to equal size), then the `fuse_coord_gloss_logits` of all shards are
all-gathered over NVLink (NCCL).  Time per step = CUDA events on each rank around {encoder + gather}, L2 flushed
between steps, max over ranks; frames/s counts the real (unpadded) sequences.  Rank 0 prints a markdown table."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from scattennet_b200 import MSCAEncoder, synth
from scattennet_b200.config import model_config
from scattennet_b200.distributed import gather_logits

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="8,16,32,64,128,256,512,1024")
ap.add_argument("--T", type=int, default=200)
ap.add_argument("--precision", default="fp16x3")
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = model_config("phoenix-2014t")
model = MSCAEncoder(cfg, 1120, precision=a.precision, use_graph=True).eval()
synth.load_synth_(model, 0)
model = model.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if rank == 0:
    print(f"### global batch sharded over {world} x B200 (T={a.T}, {a.precision}), logits all-gathered per step\n")
    print("| global batch | per GPU | ms/step (max over ranks) | frames/s (all GPUs) | gathered MB |")
    print("|---|---|---|---|---|", flush=True)
for gb in [int(x) for x in a.batches.split(",")]:
    per = -(-gb // world)  # padded shard size
    kp, mask = synth.synth_batch(per, a.T, seed=1 + rank)
    kp, mask = kp.to(dev), mask.to(dev)

    def step():
        out = model(kp, mask)
        return gather_logits(out["fuse_coord_gloss_logits"]) if world > 1 else out["fuse_coord_gloss_logits"]

    with torch.no_grad():
        for _ in range(3):
            g = step()
        steps = a.steps if per <= 64 else max(5, a.steps // 4)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for e0, e1 in ev:
            flush.zero_(); e0.record(); g = step(); e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in ev) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"| {gb} | {per} | {ms.item():.3f} | {gb * a.T / ms.item() * 1e3:,.0f} | {g.numel() * 4 / 1e6:.1f} |", flush=True)
    model._graphs.clear()
    del kp, mask, g
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
