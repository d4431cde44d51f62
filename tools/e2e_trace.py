"""simulation() end to end with DISIMPY_B200_TRACE marks (development tool, run under gpurun, or
under torchrun with N ranks): best of a few calls of the bench's sphere workload, the trace of the
last one on stderr.  PART_FIRST=<walkers> overrides the size of the first parts."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from disimpy_b200 import gradients, simulations, substrates

n, n_t = 1_000_000 * world, int(os.environ.get("NT", 10000))
g, dt = gradients.pgse(10e-3, 30e-3, n_t, [1e9], [[1.0, 0.0, 0.0]])
sub = substrates.sphere(10e-6)
for first in [int(x) for x in os.environ.get("PART_FIRST", "16384,131072").split(",")]:
    os.environ["DISIMPY_B200_PART_FIRST"] = str(first)
    os.environ.pop("DISIMPY_B200_TRACE", None)
    simulations.simulation(n, 2e-9, g, dt, sub, quiet=True)
    times = []
    for rep in range(4):
        if rep == 3:
            os.environ["DISIMPY_B200_TRACE"] = "1"
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        simulations.simulation(n, 2e-9, g, dt, sub, quiet=True)
        times.append(time.perf_counter() - t0)
    print("rank %d first part %d: simulation() %s ms" % (rank, first, [round(1e3 * t, 2) for t in times]), flush=True)
if world > 1:
    dist.destroy_process_group()
