"""Row-sharded index at the corpus sizes of BASELINE configs[3] (development tool; launch with torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/index_sharded_bench.py N [nq] [E] [K]
Every rank owns N/G corpus rows (generated on its GPU), scores all nq queries against its shard (tt_index_topk with global row
indices), the per-shard top-K lists are all-gathered over NCCL and merged on the device (tt_topk_merge).  Device-timed, max over ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import torch
import torch.distributed as dist
from pkg import _native as N
from pkg.modelling.distributed import merge_shard_results, shard_bounds

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
E = int(sys.argv[3]) if len(sys.argv) > 3 else 64
K = int(sys.argv[4]) if len(sys.argv) > 4 else 100
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = N.load()
lo, hi = shard_bounds(n_total, rank, world)
n = hi - lo
g = torch.Generator(device="cuda").manual_seed(100 + rank)
C = torch.empty((n, E), device="cuda")
for a in range(0, n, 1 << 22):
    b = min(n, a + (1 << 22))
    C[a:b] = torch.randn((b - a, E), device="cuda", generator=g).abs_() * 0.1
gq = torch.Generator(device="cuda").manual_seed(7)                       # the same queries on every rank
Q = torch.relu(torch.randn(nq, E, device="cuda", generator=gq) * 0.3)
rows_pad = ((n + 255) // 256 + 1) * 256
C32 = torch.empty_like(C); norms = torch.zeros(2 * rows_pad + rows_pad // 32 + 32, device="cuda")
st = N.stream_ptr()
N.check(lib.tt_index_prepare(C.data_ptr(), E, n, E, C32.data_ptr(), norms.data_ptr(), st))
s = torch.empty(nq, K, device="cuda"); i = torch.empty(nq, K, dtype=torch.int32, device="cuda")
ws = torch.empty(int(lib.tt_index_workspace_bytes(nq, n, E, K, N.TT_IMPL_TC, 1)), dtype=torch.uint8, device="cuda")


def run():
    N.check(lib.tt_index_topk(Q.data_ptr(), E, C.data_ptr(), E, C32.data_ptr(), norms.data_ptr(), nq, n, E, K, lo, s.data_ptr(), i.data_ptr(),
                              ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))
    return merge_shard_results(s, i, K) if world > 1 else (s, i)


for _ in range(3):
    ms_, mi_ = run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ms_, mi_ = run()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    chk = mi_.clone(); dist.broadcast(chk, src=0)
    assert torch.equal(chk, mi_), "ranks disagree on the merged result"
if rank == 0:
    owners = torch.bincount((mi_.long().clamp(min=0) // ((n_total + world - 1) // world)).reshape(-1), minlength=world).tolist()
    print(f"row-sharded index: N={n_total} over {world} GPU(s), nq={nq} E={E} K={K}: {float(t):.3f} ms/batch, {nq / float(t) * 1e3:.0f} queries/s"
          f" (top-K entries per shard: {owners})")
if world > 1:
    dist.destroy_process_group()
