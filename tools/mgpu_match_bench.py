"""Sharded 100k x 100k matcher under torchrun: nm_mgpu_match_f32 for every query-group count Q that divides the world
(world = Q x D), phase times (max over ranks) and the index hash (development aid; bench.py is the contract)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import mgpu, synth  # noqa: E402
from niftymatch_b200.dist import shard_bounds  # noqa: E402

nq = ndb = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
Bh = synth.descriptors(ndb, 2)
Ah = synth.descriptors(nq, 1, planted_from=Bh)
uid_t = torch.tensor(list(mgpu.unique_id() if rank == 0 else bytes(mgpu.ID_BYTES)), dtype=torch.uint8, device="cuda")
dist.broadcast(uid_t, 0)
mg = mgpu.MultiGpu(rank=rank, world=world, uid=bytes(uid_t.cpu().tolist()))
A = torch.from_numpy(Ah).cuda()
io = torch.full((nq,), -1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
for Q in [q for q in (1, 2, 4, 8) if world % q == 0]:
    D = world // Q
    lo, hi = shard_bounds(ndb, D, rank % D)
    Bs = torch.from_numpy(np.ascontiguousarray(Bh[lo:hi])).cuda()
    mg.set_query_groups(Q)
    mg.set_trace(False)
    for _ in range(3):
        mg.match([A], [Bs], [lo], 0.8, match_io=[io], streams=[st])
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m = mg.match([A], [Bs], [lo], 0.8, match_io=[io], streams=[st])[0]
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mg.set_trace(True)
    mg.match([A], [Bs], [lo], 0.8, match_io=[io], streams=[st])
    torch.cuda.synchronize()
    ph = mg.match_phase_ms()
    p = torch.tensor([ph["shard_scan"], ph["all_gather"], ph["merge"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(p, op=dist.ReduceOp.MAX)
    mh = m.cpu().numpy()
    h = int(np.bitwise_xor.reduce((mh.astype(np.int64) + 2) * (np.arange(nq, dtype=np.int64) * 2654435761 % (1 << 31))))
    if rank == 0:
        print(f"world {world} Q x D = {Q} x {D}: {t.item():.3f} ms  {nq * ndb / t.item() / 1e6:.0f} Gpairs/s  phases scan {p[0].item():.3f} "
              f"gather {p[1].item():.3f} merge {p[2].item():.3f}  matched {(mh >= 0).sum()} hash {h}", flush=True)
    del Bs
mg.close()
dist.destroy_process_group()
