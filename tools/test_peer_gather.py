"""torchrun check of the peer-memory logits gather against NCCL (same values, timing of both), dev tool:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/test_peer_gather.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from scattennet_b200.distributed import gather_logits, gather_logits_peer, _peer_gathers

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = (8, 50, 1120)
ok = True
for it in range(7):
    g = torch.Generator(device="cpu").manual_seed(100 * it + rank)
    x = torch.randn(*shape, generator=g).to(dev)
    a = gather_logits(x)
    b = gather_logits_peer(x)
    torch.cuda.synchronize()
    same = bool(torch.equal(a, b))
    ok = ok and same
    if not same and rank == 0:
        print("MISMATCH at iteration", it, float((a - b).abs().max()))
used_peer = any(v is not False for v in _peer_gathers.values())
x = torch.randn(*shape, device=dev)
def timeit(fn, n=50):
    for _ in range(5):
        fn(x)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn(x)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
t_nccl, t_peer = timeit(gather_logits), timeit(gather_logits_peer)
if rank == 0:
    print(f"world={world} peer path used={used_peer} equal={ok}  NCCL all-gather {t_nccl:.1f} us/call, peer push {t_peer:.1f} us/call (back to back, max over ranks)")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
