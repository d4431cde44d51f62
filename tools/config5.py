"""BASELINE config 5 through simulation(): 1e6-triangle periodic tube lattice, 180 waveforms
(60 directions x 3 shells), 1000 steps, init_pos='extra' (development / measurement tool).

    python tools/config5.py                       one GPU's share of the 8-GPU job: 1.25e7 walkers
    CONFIG5_N=100000000 torchrun ... tools/config5.py      the whole job on the ranks there are

Prints one JSON line per run (rank 0): wall time of the simulation() call (mesh sampler, walk,
reduction; the substrate is built before), the signal of the first direction of
every shell, and the containment check of SURVEY 8d config 4/5 (every final position is outside
every tube).  DISIMPY_B200_LOWRANK=0 gives the general path (18 GB of phases for 1.25e7 walkers).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["DISIMPY_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
from disimpy_b200 import gradients, meshgen, simulations, substrates  # noqa: E402

n = int(os.environ.get("CONFIG5_N", 12_500_000 * world))
n_t = int(os.environ.get("CONFIG5_NT", 1000))
radius, pitch = 5e-6, 12e-6
v, f, pad, _ = meshgen.tube_lattice(16, 16, radius, pitch, 40e-6, 128, 16)
t0 = time.perf_counter()
sub = substrates.mesh(v, f, True, padding=pad, init_pos="extra", n_sv=np.array([100, 100, 50]), quiet=True)
t_mesh = time.perf_counter() - t0
dirs = meshgen.fibonacci_sphere(60)
g, dt = gradients.pgse(10e-3, 30e-3, n_t, [1e9] * 60 + [2e9] * 60 + [3e9] * 60, np.vstack([dirs, dirs, dirs]))
simulations.simulation(10_000, 2e-9, g, dt, sub, quiet=True)           # warm-up (context, caches)
for rep in range(int(os.environ.get("CONFIG5_REPS", 2))):
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sig = simulations.simulation(n, 2e-9, g, dt, sub, quiet=True)
    t_sim = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"workload": "config 5: %d triangles, n_sv 100x100x50, %d walkers x %d steps x 180 waveforms on %d GPU(s)"
                                      % (len(f), n, n_t, world),
                          "low_rank": os.environ.get("DISIMPY_B200_LOWRANK", "1") != "0",
                          "substrates_mesh_s": round(t_mesh, 3), "simulation_s": round(t_sim, 4),
                          "walker_steps_per_s": n * n_t / t_sim,
                          "signal_over_n": [float(sig[0] / n), float(sig[60] / n), float(sig[120] / n)]}), flush=True)
if os.environ.get("CONFIG5_CONTAINMENT", "1") != "0":
    sig, pos = simulations.simulation(n, 2e-9, g, dt, sub, final_pos=True, quiet=True)
    if rank == 0:
        q = np.mod(pos[:, :2], pitch) - pitch / 2
        d = np.hypot(q[:, 0], q[:, 1])
        inscribed = radius * np.cos(np.pi / 128)
        print(json.dumps({"containment": bool(np.all(d > inscribed - 1e-12)), "finite": bool(np.all(np.isfinite(pos))),
                          "min_distance_to_axis_um": float(d.min() * 1e6), "tube_inscribed_radius_um": inscribed * 1e6}),
              flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
