"""CPU suite, part 4: the N>1 orchestration (database sharding + all-gather of per-shard top-2
records + merge) on world_size 2 and 3 with the gloo backend.  The CUDA kernels cannot run
here, so the oracle's record/merge functions stand in for nm_match_top2_f32 /
nm_match_merge_top2; what is under test is niftymatch_b200.dist (sharding arithmetic,
collective layout) and the bit-identity of sharded vs unsharded results."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._util import GOLDEN


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from niftymatch_b200.dist import match_sharded, shard_bounds
    from tests._util import load_oracle
    orc = load_oracle()
    g = np.load(os.path.join(GOLDEN, "match_200x250.npz"))
    A, B, m0 = g["A"], g["B"], g["m0"]
    lo, hi = shard_bounds(len(B), world, rank)

    def top2(a, b, off):
        an, bn = a.numpy(), np.ascontiguousarray(b.numpy())
        rec = np.zeros((len(an), 4), np.float32)
        orc.lib.orc_match_top2_true(an.ctypes.data_as(C.c_void_p), len(an), bn.ctypes.data_as(C.c_void_p), len(bn),
                                    int(off), rec.ctypes.data_as(C.c_void_p))
        return torch.from_numpy(rec)

    def merge(recs, amb):
        r = np.ascontiguousarray(recs.numpy())
        m = m0.copy()
        orc.lib.orc_merge_top2(r.ctypes.data_as(C.c_void_p), r.shape[0], r.shape[1], C.c_float(amb),
                               m.ctypes.data_as(C.c_void_p))
        return torch.from_numpy(m)

    m = match_sharded(torch.from_numpy(A), torch.from_numpy(np.ascontiguousarray(B[lo:hi])), lo, 0.8,
                      top2=top2, merge=merge)
    np.save(os.path.join(outdir, f"m{rank}.npy"), m.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_match_equals_reference(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = np.load(os.path.join(GOLDEN, "match_200x250.npz"))
    for r in range(world):
        m = np.load(tmp_path / f"m{r}.npy")
        assert np.array_equal(m, g["m"]), f"rank {r}"
