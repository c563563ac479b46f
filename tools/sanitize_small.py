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
# strip-walking blur (run with NM_BLUR_STRIP_MIN=1 to force it at this size), BGRA entry, registration, preprocessing
bgra = torch.from_numpy(np.stack([np.stack([frames[f].astype(np.uint8)] * 4, axis=-1) for f in range(2)])).cuda()
sb.run_bgra(bgra)
torch.cuda.synchronize()
r = sb.results()
print("bgra", int(r["counts"][0]), "strip forced" if os.environ.get("NM_BLUR_STRIP_MIN") else "")
tall = torch.from_numpy(synth.scene(300, 700, synth.SEED_BASE + 1)).cuda()
for rad_ in (5, 13, 16):
    tp = torch.from_numpy(np.full(2 * rad_ + 1, 1.0 / (2 * rad_ + 1), np.float32)).cuda()
    S.blur(tall, tp, rad_)
g = nm.grayscale(bgra[0, :7, :9].contiguous())
u8 = nm.cast_u8(torch.from_numpy(frames[0]).cuda(), 200)
x0, y0 = r["x"][0, :n0].contiguous(), r["y"][0, :n0].contiguous()
x1, y1 = r["x"][1, :n1].contiguous(), r["y"][1, :n1].contiguous()
c = nm.align_points(x0, y0, x1, y1, m1)
for kind in (0, 1, 2):
    H, st = nm.ransac(kind, *c, 2.0, 200, seed=3)
cb = [torch.stack([v, v]).contiguous() for v in c]
Hb, stb = nm.ransac_batch(2, *cb, torch.tensor([n0, n0 // 2], dtype=torch.int32, device="cuda"), 2.0, 200, seed=3)
torch.cuda.synchronize()
print("ransac", st.cpu().tolist(), stb.cpu().tolist())
sb.close()
print("done")
