"""Small end-to-end run of every product kernel family, for compute-sanitizer (one tool per call):
SIFT on 2 frames of 256x192 (TMA blur, extrema, compaction, orientation, descriptor), the compat
operators, the exact matcher and the tensor-core matcher with its fallback."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import synth, sift as S  # noqa: E402

frames = np.stack([synth.scene(256, 192, synth.SEED_BASE), synth.scene(256, 192, synth.SEED_BASE, shift=(1.5, 0.75))])
P = nm.SiftParams(256, 192)
sb = nm.SiftBatch(P, 2, 2048)
sb.run(torch.from_numpy(frames).cuda())
torch.cuda.synchronize()
r = sb.results()
n0, n1 = int(r["counts"][0]), int(r["counts"][1])
d0, d1 = r["desc"][0, :n0].contiguous(), r["desc"][1, :n1].contiguous()
out = sb.run_host(frames)
print("sift", n0, n1, int(out["counts"][0]))
taps, rad = nm.gaussian_taps(1.6)
img = torch.from_numpy(frames[0]).cuda()
b = S.blur(img[:77, :131].contiguous(), torch.from_numpy(taps).cuda(), rad)
nm.set_engine(0)
m0 = nm.match(d0, d1, 0.8)
nm.set_engine(1)
m1 = nm.match(d0, d1, 0.8)
B = torch.from_numpy(synth.descriptors(3000, 5)).cuda()
A = torch.from_numpy(synth.descriptors(700, 6, planted_from=B.cpu().numpy())).cuda()
pr = nm.tc_probe(A, B)
nm.set_engine(-1)
torch.cuda.synchronize()
print("match", int((m0 >= 0).sum()), bool(torch.equal(m0, m1)), "fallback rows", pr["fallback_rows"])
sb.close()
print("done")
