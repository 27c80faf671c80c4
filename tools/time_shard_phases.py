#!/usr/bin/env python
"""Per-phase device time of the row-sharded DeepFM step (CUDA events between the phases of ShardedDeepFM.train_step),
rank 0's view, config 2 per GPU.  torchrun --nproc-per-node N tools/time_shard_phases.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_repos_b200 import synth  # noqa: E402
from tf_repos_b200.sharded import ShardedDeepFM  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, B, F, K, P = int(os.environ.get("VOCAB", 200_000_000)), 8192, 39, 16, 16
m = ShardedDeepFM(F, N, K, B, update_mode="exact_deferred", epoch_steps=P, device=dev)
bt = [synth.criteo_batch(B, N, F, seed=rank * 1000 + i, device=dev) for i in range(8)]
for i in range(2 * P):
    m.train_step(*bt[i % 8])
dist.barrier(); torch.cuda.synchronize()
m._ph = []
for i in range(2 * P):
    m.train_step(*bt[i % 8])
rep = m.phase_report()
if rank == 0:
    print(json.dumps({"n_gpus": world, "vocab": N, "ms_per_step_by_phase": {k: round(v, 4) for k, v in rep.items()},
                      "sum_ms": round(sum(rep.values()), 4)}))
dist.destroy_process_group()
