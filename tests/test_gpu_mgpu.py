"""Multi-GPU C-ABI (include/nm_b200_mgpu.h) on real devices: needs >= 2 GPUs (`gpurun --gpus 2`), skipped below.
Everything is checked for BIT-IDENTITY against the single-GPU entry points."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from niftymatch_b200 import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import niftymatch_b200 as nm
    from niftymatch_b200 import mgpu
    return nm, mgpu


def _sets(nq, ndb):
    B = synth.descriptors(ndb, 2)
    A = synth.descriptors(nq, 1, planted_from=B)
    A[5] = B[7]; B[9] = B[7]                    # min1 == min2 == 0: entry stays untouched (match.cu:107)
    return A, B


@pytest.mark.parametrize("nq,ndb,cut", [(3000, 5000, 2500), (3000, 5000, 4999), (700, 900, 0), (20000, 30000, 11111)])
def test_sharded_match_two_devices_one_process(two_gpus, nq, ndb, cut):
    """nm_mgpu_match_f32 with the database split at `cut` over two devices (one all-gather over NVLink) against
    nm_match_f32 on one device: identical indices, including the untouched-entry rule and an empty shard (cut = 0)."""
    nm, mgpu = two_gpus
    A, B = _sets(nq, ndb)
    m0 = np.full(nq, 77, np.int32)
    torch.cuda.set_device(0)
    single = nm.match(torch.from_numpy(A).cuda(0), torch.from_numpy(B).cuda(0), 0.8, match_io=torch.from_numpy(m0.copy()).cuda(0))
    mg = mgpu.MultiGpu(n_dev=2)
    Ad = [torch.from_numpy(A).cuda(d) for d in range(2)]
    Bd = [torch.from_numpy(np.ascontiguousarray(B[:cut])).cuda(0), torch.from_numpy(np.ascontiguousarray(B[cut:])).cuda(1)]
    io = [torch.from_numpy(m0.copy()).cuda(d) for d in range(2)]
    for rep in range(2):                                       # second call reuses the workspaces
        out = mg.match(Ad, Bd, [0, cut], 0.8, match_io=io)
        for d in range(2):
            assert np.array_equal(out[d].cpu().numpy(), single.cpu().numpy()), (rep, d)
    assert (single.cpu().numpy() == 77)[5]
    mg.close()
    torch.cuda.set_device(0)


def test_frame_sharded_sift_two_devices_one_process(two_gpus):
    """nm_mgpu_sift_run_host splits 9 frames 5 + 4 over two devices: counts, descriptors and coordinates bitwise
    those of one device running the whole batch."""
    nm, mgpu = two_gpus
    n, w, h, cap = 9, 320, 240, 2048
    frames = np.stack([synth.scene(w, h, synth.SEED_BASE + 60 + i) for i in range(n)])
    torch.cuda.set_device(0)
    P = nm.SiftParams(w, h)
    sb = nm.SiftBatch(P, n, cap)
    ref = sb.run_host(torch.from_numpy(frames).pin_memory())
    rc, rd, rx = ref["counts"].numpy().copy(), ref["desc"].numpy().copy(), ref["x"].numpy().copy()
    sb.close()
    mg = mgpu.MultiGpu(n_dev=2)
    mg.sift_create(P, n, cap)
    out = mg.sift_run_host(torch.from_numpy(frames).pin_memory())
    assert np.array_equal(out["counts"].numpy(), rc[:n]) and rc.min() > 50
    for f in range(n):
        k = int(rc[f])
        assert np.array_equal(out["desc"][f, :k].numpy(), rd[f, :k]), f
        assert np.array_equal(out["x"][f, :k].numpy(), rx[f, :k]), f
    mg.close()
    torch.cuda.set_device(0)


_RANK_SCRIPT = r"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, sys.argv[1])
rank, world, idfile, outfile = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
torch.cuda.set_device(rank)
import niftymatch_b200 as nm
from niftymatch_b200 import mgpu, synth
from niftymatch_b200.dist import shard_bounds
if rank == 0:
    uid = mgpu.unique_id()
    open(idfile + '.tmp', 'wb').write(uid); os.replace(idfile + '.tmp', idfile)
else:
    while not os.path.exists(idfile): time.sleep(0.05)
    uid = open(idfile, 'rb').read()
mg = mgpu.MultiGpu(rank=rank, world=world, uid=uid)
B = synth.descriptors(9000, 2); A = synth.descriptors(4000, 1, planted_from=B)
lo, hi = shard_bounds(9000, world, rank)
Ad = torch.from_numpy(A).cuda(); Bd = torch.from_numpy(np.ascontiguousarray(B[lo:hi])).cuda()
st = torch.cuda.current_stream()
for _ in range(3):
    out = mg.match([Ad], [Bd], [lo], 0.8, streams=[st])
torch.cuda.synchronize()
np.save(outfile, out[0].cpu().numpy())
mg.close()
"""


def test_sharded_match_one_process_per_gpu(two_gpus):
    """The torchrun / MPI model: two processes, one GPU each, nm_mgpu_create_rank from a unique id passed through a
    file, caller streams (nothing synchronises inside).  Both ranks must hold the single-GPU result."""
    nm, mgpu = two_gpus
    B = synth.descriptors(9000, 2)
    A = synth.descriptors(4000, 1, planted_from=B)
    torch.cuda.set_device(0)
    single = nm.match(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), 0.8).cpu().numpy()
    with tempfile.TemporaryDirectory() as td:
        idf = os.path.join(td, "uid")
        procs = [subprocess.Popen([sys.executable, "-c", _RANK_SCRIPT, ROOT, str(r), "2", idf, os.path.join(td, f"m{r}.npy")],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
        for p in procs:
            try:
                out, err = p.communicate(timeout=240)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                pytest.fail("rank processes timed out")
            assert p.returncode == 0, err[-3000:]
        for r in range(2):
            assert np.array_equal(np.load(os.path.join(td, f"m{r}.npy")), single), r
    assert (single >= 0).sum() > 400


@pytest.mark.parametrize("nq", [4000, 4001])
def test_query_groups_two_devices(two_gpus, nq):
    """Two-dimensional sharding, Q = 2 query groups x D = 1 database shards on two devices: every device scans its
    half of the queries against the whole database; one all-gather, one merge per query block (the last block is one
    row short for an odd query count).  Same indices as the single-GPU call."""
    nm, mgpu = two_gpus
    A, B = _sets(nq, 6000)
    torch.cuda.set_device(0)
    single = nm.match(torch.from_numpy(A).cuda(0), torch.from_numpy(B).cuda(0), 0.8).cpu().numpy()
    mg = mgpu.MultiGpu(n_dev=2)
    mg.set_query_groups(2)
    out = mg.match([torch.from_numpy(A).cuda(d) for d in range(2)], [torch.from_numpy(B).cuda(d) for d in range(2)], [0, 0], 0.8)
    for d in range(2):
        assert np.array_equal(out[d].cpu().numpy(), single), d
    mg.close()
    torch.cuda.set_device(0)
