"""One eager encoder step (with the BiLSTM alignment head) in which the FIRST launch of every distinct kernel /
GEMM shape is bracketed by cudaProfilerStart/Stop, for `ncu --set full --profile-from-start off` (dev tool; the
report of a whole step would be ~340 MB, this keeps it to ~14 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import MSCAEncoder, synth
from scattennet_b200 import functional as F_
from scattennet_b200.config import model_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
MAX_LINEAR = 7
cfg = model_config("phoenix-2014t")
m = MSCAEncoder(cfg, 1120, precision="fp16x3", alignment=True).eval()
synth.load_synth_(m, 0)
m = m.cuda()
kp, mask = synth.synth_batch(B, T, seed=1)
kp, mask = kp.cuda(), mask.cuda()

seen, picked, armed = set(), [], [False]
orig_enter, orig_exit = F_._timed.__enter__, F_._timed.__exit__


def enter(self):
    key = (self.name, self.flops)
    n_lin = sum(1 for k in seen if k[0] == "linear_tc_kernel")
    if armed[0] and key not in seen and (self.name != "linear_tc_kernel" or n_lin < MAX_LINEAR):
        seen.add(key)
        picked.append(key)
        self._prof = True
        torch.cuda.profiler.start()
    else:
        self._prof = False
    return orig_enter(self)


def exit_(self, *exc):
    r = orig_exit(self, *exc)
    if self._prof:
        torch.cuda.profiler.stop()
    return r


F_._timed.__slots__  # noqa: B018  (slots class: attach the flag through a subclass-free dict below)
with torch.no_grad():
    for _ in range(2):
        m(kp, mask)
    torch.cuda.synchronize()
    class Timed(F_._timed):
        __slots__ = ("_prof",)
        __enter__ = enter
        __exit__ = exit_
    F_._timed = Timed
    armed[0] = True
    out = m(kp, mask)
    torch.cuda.synchronize()
print("profiled", len(picked), "launches:")
for k in picked:
    print("  ", k)
