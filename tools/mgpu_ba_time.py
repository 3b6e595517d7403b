"""Per-class device time of the sharded c4 BA under torchrun (every rank prints its own timers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions
prob, _ = synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                              seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)
opt = BaOptions.defaults(); opt.device = local
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.from_numpy(api.nccl_unique_id()))
dist.broadcast(uid, 0)
comm = api.BaComm(uid.cpu().numpy(), rank, world, local)
shard, _ = prob.shard(rank, world)
s = api.BaSolver(shard, opt, comm)
s.run_fixed(3); s.reset(); s.profile_enable(True)
info = s.run_fixed(8)
print(f"[rank {rank}/{world}] cost {info[1]:.9g}", {k: round(1e3 * v[0] / max(1, v[1]), 1) for k, v in s.timers().items()}, file=sys.stderr, flush=True)
dist.barrier(); s.close(); comm.close(); dist.destroy_process_group()
