"""torchrun --nproc-per-node N scripts/dist_check.py : sharded placement over N GPUs + one NCCL allgather
must reproduce the single-GPU result of the whole batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from usher_b200 import capi, dist as ud
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s = capi.Synth(200_000, 30.0, 30000, 0, 123)
m = capi.Mat.from_flat_struct(s.flat, device=lr)
ok = True
for fam, n in ((0, 1000), (2, 333), (1, 5)):
    sp, sc, _ = s.samples(n, fam, 7 + fam)
    full = ud.place_sharded(m, sp, sc, rank, world, device=torch.device("cuda", lr))
    single = m.place_batch(sp, sc)["placements"]
    same = all(np.array_equal(full[k], single[k]) for k in ("score", "best_node", "best_j", "num_best", "has_unique"))
    ok = ok and same
    if rank == 0:
        print(f"family {fam}: {n} samples over {world} GPUs == single GPU: {same}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("DIST_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
