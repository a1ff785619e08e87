"""2-GPU check of the universal-perturbation mode: every rank evaluates its shard, the closure all-reduces
[grad | loss] over NCCL, and all ranks must hold bit-identical deltas afterwards.  Launch with torchrun."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from pcfa_b200.adapter import build_network
from pcfa_b200.attack import UniversalAttack
from pcfa_b200.networks.weights import synthetic_pair

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net = build_network("RAFT", device=dev, seed=0, gain=0.5)
ua = UniversalAttack(net, "RAFT", (128, 160), dev, delta_bound=0.005, joint_perturbation=True, iters=4)
per_rank = 2
pairs = [synthetic_pair(rank * per_rank + i, 128, 160) for i in range(per_rank)]
i1 = torch.cat([p[0] for p in pairs]).to(dev); i2 = torch.cat([p[1] for p in pairs]).to(dev)
t0 = time.time()
stats = ua.run_batch(i1, i2, steps=2)
torch.cuda.synchronize()
gathered = [torch.empty_like(ua.delta1) for _ in range(world)]
dist.all_gather(gathered, ua.delta1.detach())
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    print("universal 2-GPU: stats", stats, "closures", ua.closure_evals, "identical deltas across ranks:", same,
          "l2", ua.l2_norms()[2], "secs %.2f" % (time.time() - t0))
assert same
dist.destroy_process_group()
