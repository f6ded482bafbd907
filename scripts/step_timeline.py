#!/usr/bin/env python
"""In-situ kernel timeline of the train step (CUDA-graph replay, warm caches) through torch.profiler / CUPTI: per-kernel mean
duration over N steps and the chronological layout of one step (start offset, duration, stream).  Development tool.
    python scripts/step_timeline.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
attach = None
if world > 1:      # torchrun: the data-parallel step (row-sharded tables, device barriers), rank 0 reports
    import torch.distributed as dist

    from pkg.modelling.distributed import DataParallel

    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    attach = lambda m: DataParallel(m)   # noqa: E731
model = bench.build_gpu_model(attach)
rng = np.random.default_rng(1 + rank)
batches = [{k: torch.from_numpy(v).cuda() for k, v in bench.make_batch(rng, B).items()} for _ in range(4)]
for i in range(6):
    model.train_step(batches[i % 4])
torch.cuda.synchronize()
steps = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        model.train_step(batches[i % 4])
    torch.cuda.synchronize()
if rank != 0:
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
print(f"{len(evs)} device activities over {steps} steps ({len(evs) / steps:.1f} per step)")
agg = {}
for e in evs:
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.elapsed_us()
print(f"{'kernel':72s} {'n/step':>6s} {'us each':>8s} {'us/step':>8s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n / steps:6.1f} {us / n:8.2f} {us / steps:8.2f}")
# one step in the middle, chronologically
per = len(evs) // steps
mid = evs[per * (steps // 2): per * (steps // 2 + 1)]
t0 = mid[0].time_range.start
print("\none step (offset us, duration us, stream):")
for e in mid:
    print(f"  {e.time_range.start - t0:8.1f} {e.time_range.elapsed_us():7.1f}  s{getattr(e, 'device_index', 0)}:{getattr(e, 'stream', '?')}  {e.name[:80]}")
print(f"  step span {mid[-1].time_range.end - t0:.1f} us")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
