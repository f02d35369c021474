"""Multi-GPU check of the prototype exchange (run under torchrun, one rank per GPU):
NVLink peer-memory exchange == NCCL all-gather exchange == single-process whole-batch prototypes, bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from wseg_b200 import ops, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n_local, g = 4, 16
ok = True
for it in range(6):                                   # several steps: exercises the parity double buffer and the step counter
    full = synth.head_outputs(n_local * world, (g * 8, g * 8), seed=100 + it)
    label = synth.label_with_bg(synth.labels(n_local * world, seed=it)).to(dev)
    norm_full, _ = ops.seed_labels(full["cam_rv_down"].to(dev), label, 0.2)
    fp_full = full["f_proj"].to(dev)
    sl = slice(rank * n_local, (rank + 1) * n_local)
    for view in (0, 1):
        p_nv, i_nv = ops.prototypes(norm_full[sl].contiguous(), fp_full[sl].contiguous(), group=dist.group.WORLD, view=view,
                                    exchange="nvlink")
        p_nc, i_nc = ops.prototypes(norm_full[sl].contiguous(), fp_full[sl].contiguous(), group=dist.group.WORLD, view=view,
                                    exchange="nccl")
        k = g * g // 8                                                   # single-process reference on the whole batch
        p_1, i_1 = ops.proto_finalize(*ops.proto_topk_local(norm_full, fp_full, k), k)
        torch.cuda.synchronize()
        good = torch.equal(p_nv, p_nc) and torch.equal(i_nv, i_nc) and torch.equal(i_nv, i_1) and torch.equal(p_nv, p_1)
        ok = ok and good
        if not good:
            print(f"rank {rank} it {it} view {view}: nv==nc {torch.equal(p_nv, p_nc)} idx {torch.equal(i_nv, i_nc)} "
                  f"vs single {torch.equal(p_nv, p_1)} maxdiff {(p_nv - p_1).abs().max().item():.3e}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("EXCHANGE PARITY", "OK" if int(flag) == 1 else "FAILED", f"(world {world}, symmetric path active: {any(v is not None for v in ops._SYMM_MODE.values())})", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if int(flag) == 1 else 1)
