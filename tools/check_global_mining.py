"""Multi-GPU check of the GLOBAL hard-pixel mining (run under torchrun, one rank per GPU):
for every rank, mined sets and feature gradients of its shard == the single-process computation on the whole batch, and
the mean over ranks of the reported losses == the single-process losses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from wseg_b200 import loss as L, ops, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n_local, g = 4, 16
# group=None means "the default group" once torch.distributed is initialised: the single-process reference needs a group of one
solo = [dist.new_group([r]) for r in range(world)][rank]
ok = True
for it in range(5):                                   # several steps: exercises the parity double buffer and the step counter
    n = n_local * world
    a = {k: v.to(dev) for k, v in synth.head_outputs(n, (g * 8, g * 8), seed=200 + it).items()}
    b = {k: v.to(dev) for k, v in synth.head_outputs(n, (g * 8, g * 8), seed=300 + it).items()}
    label = synth.label_with_bg(synth.labels(n, seed=it)).to(dev)
    P = n * g * g
    k1, k2 = ops.mining_keys(P, 9, dev, step=it, stream_id=1), ops.mining_keys(P, 9, dev, step=it, stream_id=2)
    # single process, whole batch
    one = L.contrast_terms(a["f_proj"], a["cam_rv_down"], b["f_proj"], b["cam_rv_down"], label, 0.2, grid=(g, g),
                           keys=(k1, k2), want_mined=True, group=solo)
    # this rank's shard, global prototypes (group) + global mining
    sl = slice(rank * n_local, (rank + 1) * n_local)
    ps = slice(rank * n_local * g * g, (rank + 1) * n_local * g * g)
    mine = L.contrast_terms(a["f_proj"][sl].contiguous(), a["cam_rv_down"][sl].contiguous(), b["f_proj"][sl].contiguous(),
                            b["cam_rv_down"][sl].contiguous(), label[sl].contiguous(), 0.2, grid=(g, g),
                            keys=(k1[ps].contiguous(), k2[ps].contiguous()), want_mined=True, group=dist.group.WORLD,
                            mining="global")
    torch.cuda.synchronize()
    good = True
    for v in "12":
        good &= torch.equal(mine["mined" + v], one["mined" + v][ps])
        good &= torch.equal(mine["prototypes" + v], one["prototypes" + v])
        # gradients: the single-process loss is a mean over R times more pixels / the global intra loss, and DDP averages
        # the ranks' gradients: grad_single == grad_rank / R
        gs, gr = one["grad_f" + v][sl], mine["grad_f" + v] / world
        good &= bool(((gs - gr).abs().max() <= 1e-6 * gs.abs().max().clamp_min(1e-30)).item())
    losses = mine["losses_raw"].clone()
    dist.all_reduce(losses, op=dist.ReduceOp.SUM)
    losses /= world
    good &= bool(torch.allclose(losses[6:10], one["losses_raw"][6:10], rtol=2e-5, atol=0))
    ok = ok and good
    if not good:
        print(f"rank {rank} it {it}: proto1 {torch.equal(mine['prototypes1'], one['prototypes1'])} proto2 "
              f"{torch.equal(mine['prototypes2'], one['prototypes2'])} pl1 {torch.equal(mine['pseudo_label1'], one['pseudo_label1'][ps])} "
              f"raw {mine['losses_raw'][:6].tolist()} one {one['losses_raw'][:6].tolist()}", flush=True)
        print(f"rank {rank} it {it}: mined1 {torch.equal(mine['mined1'], one['mined1'][ps])} "
              f"mined2 {torch.equal(mine['mined2'], one['mined2'][ps])} "
              f"losses {losses[6:10].tolist()} vs {one['losses_raw'][6:10].tolist()} "
              f"grad rel {((one['grad_f1'][sl] - mine['grad_f1'] / world).abs().max() / one['grad_f1'][sl].abs().max()).item():.2e}",
              flush=True)
    # and the shard-local mining differs (sanity: the global path is not a no-op)
    if it == 0:
        shard = L.contrast_terms(a["f_proj"][sl].contiguous(), a["cam_rv_down"][sl].contiguous(), b["f_proj"][sl].contiguous(),
                                 b["cam_rv_down"][sl].contiguous(), label[sl].contiguous(), 0.2, grid=(g, g),
                                 keys=(k1[ps].contiguous(), k2[ps].contiguous()), want_mined=True, group=dist.group.WORLD)
        differs = not torch.equal(shard["mined1"], mine["mined1"])
        if rank == 0:
            print("shard-local mining differs from global:", differs, flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("GLOBAL MINING PARITY", "OK" if int(flag) == 1 else "FAILED", f"(world {world})", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if int(flag) == 1 else 1)
