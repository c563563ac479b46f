"""One 100k x 100k (or given size) match through the C-ABI, for ncu.  Usage: match_once.py [nA nB engine reps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import synth  # noqa: E402

nA = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nB = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
eng = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
Bh = synth.descriptors(nB, 2)
Ah = synth.descriptors(nA, 1, planted_from=Bh)
A, B = torch.from_numpy(Ah).cuda(), torch.from_numpy(Bh).cuda()
nm.set_engine(eng)
for _ in range(reps):
    m = nm.match(A, B, 0.8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m = nm.match(A, B, 0.8)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"engine {eng} {nA}x{nB}: {ms:.3f} ms {nA * nB / ms / 1e6:.1f} Gpairs/s matched {int((m >= 0).sum())}")
