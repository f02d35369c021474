"""When each milestone of the resident step is reached inside the replayed CUDA graph (%globaltimer stamps written by
one-thread kernels on the chains' own streams).  usage: python tools/step_timeline.py [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
hp = bench.HotPath(bench.N_PER_GPU, dev, seed=1234)
hp.marks = {"names": {}, "buf": torch.zeros(64, dtype=torch.int64, device=dev)}
stepper = bench.GraphedStep(hp)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
acc = []
for it in range(steps + 3):
    flush.zero_()
    stepper.step()
    torch.cuda.synchronize(dev)
    if it >= 3:
        acc.append(hp.marks["buf"].cpu().clone())
a = torch.stack(acc)
names = sorted(hp.marks["names"].items(), key=lambda kv: kv[1])
t0 = a[:, hp.marks["names"]["start"]]
print(f"median over {steps} replays, us after the step's first stamp ({stepper.launches} launches per step incl. the stamps)")
for n, i in sorted(names, key=lambda kv: float(((a[:, kv[1]] - t0).double() / 1e3).median())):
    x = (a[:, i] - t0).double() / 1e3
    print(f"{x.median():8.1f}  (min {x.min():7.1f} max {x.max():7.1f})  {n}")
