"""Diagnostic: descriptors of a 16-frame 1080p batch (two scenes alternating) against the CPU oracle."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import synth  # noqa: E402
from tests._util import load_oracle, ang_diff  # noqa: E402

orc = load_oracle()
big = [synth.scene(1920, 1080, synth.SEED_BASE), synth.scene(1920, 1080, synth.SEED_BASE, shift=(2.5, 1.25))]
P2 = nm.SiftParams(1920, 1080)
for nb in (16, 2):
    sb2 = nm.SiftBatch(P2, nb, 16384)
    sb2.run(torch.from_numpy(np.stack([big[i % 2] for i in range(nb)])).cuda())
    torch.cuda.synchronize()
    r2 = sb2.results()
    for f in (0, 1, nb - 1):
        c = orc.sift_frame(big[f % 2], want_levels=False, capacity=16384)
        n = int(r2["counts"][f].item())
        d = r2["desc"][f, :n].cpu().numpy()
        o = r2["orient"][f, :n].cpu().numpy()
        rel = np.linalg.norm(d - c["desc"], axis=1) / np.maximum(np.linalg.norm(c["desc"], axis=1), 1e-30)
        bad = np.nonzero(rel > 2e-5)[0]
        print(f"batch {nb} frame {f}: n {n} oracle {c['n']} kpts equal {np.array_equal(r2['kpts'][f, :n].cpu().numpy(), c['kpts'])} "
              f"max rel {rel.max():.3g} bad {len(bad)}")
        for i in bad[:5]:
            print("   idx", i, "rel", rel[i], "|c|", np.linalg.norm(c["desc"][i]), "|d|", np.linalg.norm(d[i]), "orient p", o[i], "o", c["orient"][i],
                  "kp", c["kpts"][i])
    sb2.close()
