"""Multi-GPU check of simulation() (run under torchrun on >= 2 GPUs, e.g.
    gpurun --gpus 2 -- 'python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py'
): signals, final positions and per-walker signals of the sharded run must equal the single-GPU
run of the same call -- positions bit for bit, signals to summation-order rounding -- for the
pipelined path with round-robin parts, for a small run (contiguous shards), a many-measurement
protocol, a mesh, and a run with fewer walkers than ranks.  Then rank 0 alone repeats the calls as a
single process driving all the GPUs (the device list of SURVEY.md 8b)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["DISIMPY_B200_DEVICE"] = str(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
from disimpy_b200 import gradients, meshgen, simulations, substrates  # noqa: E402

g1, dt1 = gradients.pgse(10e-3, 30e-3, 200, [1e9, 2e9], [[1.0, 0, 0], [0, 0.6, 0.8]])
g12, dt12 = gradients.pgse(10e-3, 30e-3, 100, np.linspace(5e8, 3e9, 12), np.random.RandomState(1).normal(size=(12, 3)))
v, f, pad, _ = meshgen.tube_lattice(2, 2, 2e-6, 5e-6, 6e-6, 16, 3)
mesh = substrates.mesh(v, f, True, padding=pad, init_pos="extra", n_sv=np.array([8, 8, 6]), quiet=True)
cases = [("sphere 600k (round-robin parts)", 600_000, g1, dt1, substrates.sphere(5e-6)),
         ("cylinder 300k", 300_000, g1, dt1, substrates.cylinder(3e-6, np.array([0.2, 1.0, -0.3]))),
         ("sphere 1000 (contiguous shards)", 1000, g1, dt1, substrates.sphere(5e-6)),
         ("ellipsoid 400k, 12 measurements", 400_000, g12, dt12, substrates.ellipsoid(np.array([5e-6, 3e-6, 2e-6]))),
         ("periodic mesh 50k, init_pos extra", 50_000, g1, dt1, mesh)]
simulations._SHARDED_FILL_MIN = 0     # the mesh sampler's threads are dealt to the ranks even for this small run
mesh_intra = substrates.mesh(v, f, True, padding=pad, init_pos="intra", n_sv=np.array([8, 8, 6]), quiet=True)
cases.append(("periodic mesh 30001, init_pos intra", 30_001, g1, dt1, mesh_intra))
cases.append(("sphere, 1 walker (fewer walkers than ranks: empty shards)", 1, g1, dt1, substrates.sphere(5e-6)))
real_dist = simulations._dist
for name, n, g, dt, sub in cases:
    t0 = time.time()
    sig, pos = simulations.simulation(n, 2e-9, g, dt, sub, seed=5, final_pos=True, quiet=True)
    t_multi = time.time() - t0
    allsig = simulations.simulation(n, 2e-9, g, dt, sub, seed=5, all_signals=True, quiet=True)
    simulations._dist = lambda: (0, 1, None)          # the same calls on this GPU alone
    sig1, pos1 = simulations.simulation(n, 2e-9, g, dt, sub, seed=5, final_pos=True, quiet=True)
    allsig1 = simulations.simulation(n, 2e-9, g, dt, sub, seed=5, all_signals=True, quiet=True)
    simulations._dist = real_dist
    assert np.array_equal(pos, pos1), name
    assert np.allclose(sig, sig1, rtol=1e-12, atol=0), name
    assert np.allclose(allsig, allsig1, rtol=0, atol=1e-9), name
    if rank == 0:
        print("ok  %-40s %d ranks, %.0f ms" % (name, world, 1e3 * t_multi), flush=True)
dist.barrier()
dist.destroy_process_group()

# The device list of ONE process (a plain script on a multi-GPU box, no torchrun): rank 0 alone drives
# all the GPUs of the job and must get what one GPU gives.
if rank == 0:
    os.environ.pop("DISIMPY_B200_DEVICE", None)
    os.environ["DISIMPY_B200_MIN_WALKERS_PER_DEVICE"] = "1000"
    for name, n, g, dt, sub in cases[:-1]:
        out = {}
        for devs in (",".join(str(d) for d in range(world)), "0"):
            os.environ["DISIMPY_B200_DEVICES"] = devs
            t0 = time.time()
            sig, pos = simulations.simulation(n, 2e-9, g, dt, sub, seed=5, final_pos=True, quiet=True)
            out[devs] = (sig, pos, time.time() - t0)
        (sig, pos, t_multi), (sig1, pos1, t_one) = out.values()
        assert np.array_equal(pos, pos1), name
        assert np.allclose(sig, sig1, rtol=1e-12, atol=0), name
        print("ok  %-40s one process, devices %s: %.0f ms (device 0 alone: %.0f ms)"
              % (name, list(out)[0], 1e3 * t_multi, 1e3 * t_one), flush=True)
