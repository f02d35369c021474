"""Per-rank timeline of the resident step under torchrun: when each rank's step starts and ends, when it starts waiting for
its peers' prototype pushes and when it sees each peer's flag (%globaltimer, one clock for the GPUs of a box).
usage: torchrun --nproc-per-node N tools/scale_timeline.py [steps]   -> one block of lines per rank, rank 0 prints the summary"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from wseg_b200 import _lib, ops

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
hp = bench.HotPath(bench.N_PER_GPU, dev, seed=1234 + rank, group=group)
stamps = torch.zeros(2 * 17, dtype=torch.int64, device=dev)
edge = torch.zeros(2, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.wsegk_debug_exchange_stamps(stamps.data_ptr())
stepper = bench.GraphedStep(hp)             # the stamps pointer is captured with the graph
lib.wsegk_debug_exchange_stamps(None)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cur = lambda: torch.cuda.current_stream(dev).cuda_stream
rows = []
for it in range(steps + 3):
    flush.zero_()
    lib.wsegk_debug_stamp(edge.data_ptr(), cur())
    stepper.step()
    lib.wsegk_debug_stamp(edge.data_ptr() + 8, cur())
    torch.cuda.synchronize(dev)
    if it >= 3:
        e, s = edge.cpu().tolist(), stamps.cpu().tolist()
        rows.append([e[0], e[1]] + [s[0]] + s[1:1 + world] + [s[17]] + s[18:18 + world])
t = torch.tensor(rows, dtype=torch.int64, device=dev)
if world > 1:
    allr = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    a = torch.stack(allr).cpu()                       # [rank, step, col]
    print(f"{world} ranks, {steps} steps; all times in us relative to the EARLIEST step start of the step (globaltimer)")
    print("rank: start  end(step us) | view 1: wait from, saw flags of ranks ... | view 2: wait from, saw ...")
    for k in range(steps):
        t0 = int(a[:, k, 0].min())
        for r in range(world):
            x = [(int(v) - t0) / 1e3 for v in a[r, k]]
            v1 = " ".join(f"{v:7.1f}" for v in x[3:3 + world])
            v2 = " ".join(f"{v:7.1f}" for v in x[4 + world:4 + 2 * world])
            print(f"step {k:2d} rank {r}: {x[0]:7.1f} {x[1]:7.1f} ({x[1] - x[0]:6.1f}) | {x[2]:7.1f}  {v1} | {x[3 + world]:7.1f}  {v2}")
    dur = (a[:, :, 1] - a[:, :, 0]).double() / 1e3
    skew = (a[:, :, 0].max(0).values - a[:, :, 0].min(0).values).double() / 1e3
    wait1 = (a[:, :, 3:3 + world].max(2).values - a[:, :, 2]).double() / 1e3
    print(f"summary: step duration mean {dur.mean():.1f} us (max over ranks, mean over steps {dur.max(0).values.mean():.1f}); start skew mean {skew.mean():.1f} us max {skew.max():.1f}; "
          f"time inside the view-1 flag wait: mean {wait1.mean():.1f} us max {wait1.max():.1f}")
if world > 1:
    del stepper
    torch.cuda.synchronize(dev)
    dist.barrier()
    os._exit(0)
