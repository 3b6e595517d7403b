"""Multi-GPU BA parity check: run under torchrun with N ranks; every rank solves its shard of the
same problem with the NCCL all-reduce path, rank 0 compares against a single-GPU solve."""
import faulthandler, os, sys
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions

def log(*a):
    print(f"[rank {rank}]", *a, file=sys.stderr, flush=True)

# points in map-creation order: every rank sees different camera pairs
prob, truth = synth.make_ba_scene(4, 30, 5000, 1280, 720, seed=21, m_con=4, n_con=3, sort_by_home=True)
opt = BaOptions.defaults()
opt.device = local
opt.outer_iters, opt.inner_iters = 2, 6
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.from_numpy(api.nccl_unique_id()))
dist.broadcast(uid, 0)
log("uid ok")
comm = api.BaComm(uid.cpu().numpy(), rank, world, local)
log("comm ok")
shard, (lo, hi) = prob.shard(rank, world)
log("shard", lo, hi, shard.nobs)
s = api.BaSolver(shard, opt, comm)
info = s.run()
s.download()
log("solved", info[:3], info[9], info[10])
# gather the points back on rank 0
Xs = [None] * world
dist.all_gather_object(Xs, (lo, hi, shard.X.copy(), shard.outlier.copy()))
if rank == 0:
    ref = prob.copy()
    o1 = BaOptions.defaults(); o1.device = local; o1.outer_iters, o1.inner_iters = 2, 6
    info1 = api.ba_solve(ref, o1)
    X = np.concatenate([x[2] for x in sorted(Xs)])
    out = np.concatenate([x[3] for x in sorted(Xs)])
    dX = np.abs(X - ref.X).max(); dR = np.abs(shard.R - ref.R).max(); dt = np.abs(shard.t - ref.t).max()
    print(f"multi-GPU ({world}) vs single GPU: max|dX| {dX:.3e} max|dR| {dR:.3e} max|dt| {dt:.3e} "
          f"cost {info[1]:.9g} vs {info1[1]:.9g} trials {info[10]} vs {info1[10]} "
          f"outliers equal {np.array_equal(out, ref.outlier)} info[13] {info[13]:.0f} vs flags {int(out.sum())}")
    assert info[13] == out.sum(), "info[13] of cosl_ba_solver_run must count the outliers of ALL ranks"
    ok = dX < 1e-6 and dR < 1e-8 and abs(info[1] - info1[1]) < 1e-8 * info1[1] and info[10] == info1[10]
    print("MGPU_PARITY_OK" if ok else "MGPU_PARITY_FAIL")
dist.barrier()
s.close(); comm.close()
dist.destroy_process_group()
