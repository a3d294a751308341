"""Dev probe: the CPU oracle timed in a fresh process under different process states (why bench.py's in-process cpu_baseline is
2x faster than the stand-alone reference arm)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
mode = sys.argv[1]
if mode == "cuda":
    torch.cuda.init(); x = torch.zeros(1, device="cuda"); torch.cuda.synchronize()
if mode == "pinned":
    x = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
if mode == "profiler":
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU]) as p:
        torch.randn(1000, 1000) @ torch.randn(1000, 1000)
if mode == "ftz":
    print("set_flush_denormal ->", torch.set_flush_denormal(True))
if mode == "gpuarm":
    from scattennet_b200 import MSCAEncoder, synth
    from scattennet_b200.config import model_config
    m = MSCAEncoder(model_config("phoenix-2014t"), 1120, precision="fp16x3", use_graph=True).eval()
    synth.load_synth_(m, 0); m = m.cuda()
    kp, mask = synth.synth_batch(8, 200, seed=1)
    with torch.no_grad():
        for _ in range(3): m(kp.cuda(), mask.cuda())
    torch.cuda.synchronize()
import bench
if mode.startswith("threads"):
    n = int(mode[7:])
    bench.host_cores = lambda: n
r = bench.time_oracle(8, 4, 60.0)
print(mode, round(r["ms_per_step"], 1), round(r["best_ms"], 1), r["cores"], torch.get_num_threads(), torch.__config__.parallel_info().split("\n")[0:3])
