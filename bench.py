#!/usr/bin/env python
"""
bench.py -- headline benchmark of the two-tower hot path (BASELINE.json: "train examples/s; index
queries/s (top-100, 105k items) at 1/2/4/8 B200").

    python bench.py --gpus 1 --steps K --warmup W                       (ours)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                (CPU arm: the oracle port on host cores)

Workload (configs[1], "c2"): H&M-shaped synthetic data -- 1 371 980 customers, 105 542 articles (Zipf), side
features age (f32), product_type (131 -> e16), colour (50 -> e8); id embeddings and joint dim 64, no hidden
layers, logQ-corrected in-batch softmax, Adagrad lr 0.05, batch 8192 per GPU (weak scaling).
One "step" = one TwoTowerModel.train_step on one batch.  `value` = examples/s with the batch ids already
resident in HBM; `e2e` = the same through the public API from pinned host buffers, loss read back each step.
The index half of the metric is reported in the same JSON line under "index".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "hm-retrieval-two-tower_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

V_CUSTOMERS, V_ARTICLES, V_PTYPE, V_COLOUR = 1_371_980, 105_542, 131, 50
E_ID, E_PTYPE, E_COLOUR, JOINT = 64, 16, 8, 64
INDEX_K, INDEX_BQ = 100, 2048
METRIC = "train examples/s; index queries/s (top-100, 105k items) at 1/2/4/8 B200"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------------
# synthetic data (seeded; pre-encoded row ids, 0 = OOV)
# ---------------------------------------------------------------------------------------------------
def zipf_articles(rng, n):
    # Zipf(s ~ 1.0) over {1..V_ARTICLES} by inverse-CDF on the harmonic weights
    w = 1.0 / np.arange(1, V_ARTICLES + 1, dtype=np.float64)
    cdf = np.cumsum(w / w.sum())
    return (np.searchsorted(cdf, rng.random(n), side="left") + 1).clip(1, V_ARTICLES).astype(np.int32)


def article_probs():
    w = 1.0 / np.arange(1, V_ARTICLES + 1, dtype=np.float64)
    return (w / w.sum()).astype(np.float32)


def make_batch(rng, b):
    art = zipf_articles(rng, b)
    return {
        "age": (np.clip(rng.normal(36, 14, size=b), 16, 99) / 100.0).astype(np.float32).reshape(b, 1),
        "customer_id": rng.integers(1, V_CUSTOMERS + 1, size=(b, 1)).astype(np.int32),
        "article_id": art.reshape(b, 1),
        "product_type_name": (art % V_PTYPE + 1).astype(np.int32).reshape(b, 1),   # functionally dependent on the article
        "colour_group_name": (art % V_COLOUR + 1).astype(np.int32).reshape(b, 1),
    }


BYTES_PER_EXAMPLE_H2D = 4 * 5


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port with all host threads
# ---------------------------------------------------------------------------------------------------
def build_cpu_model(seed=0):
    import torch
    from oracle.torch_cpu_port import CpuTwoTower

    logp = torch.zeros(V_ARTICLES + 1)
    logp[1:] = torch.from_numpy(np.log(article_probs()))
    return CpuTwoTower(q_cat=[(V_CUSTOMERS + 1, E_ID)], q_num=1,
                       c_cat=[(V_ARTICLES + 1, E_ID), (V_PTYPE + 1, E_PTYPE), (V_COLOUR + 1, E_COLOUR)], c_num=0, joint=JOINT,
                       log_p_rows=logp, lr=0.05, seed=seed)


def cpu_train_rate(batch, warmup, steps, seed=1):
    """examples/s of the torch-CPU port on `steps` batches of the same workload."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_cpu_model()
    rng = np.random.default_rng(seed)
    batches = []
    for _ in range(4):
        b = make_batch(rng, batch)
        batches.append(([torch.from_numpy(b["customer_id"].reshape(-1).astype(np.int64))], [torch.from_numpy(b["age"].reshape(-1))],
                        [torch.from_numpy(b[k].reshape(-1).astype(np.int64)) for k in ("article_id", "product_type_name", "colour_group_name")], []))
    it = [0]

    def step():
        q_ids, q_nums, c_ids, c_nums = batches[it[0] % len(batches)]
        it[0] += 1
        model.train_step(q_ids, q_nums, c_ids, c_nums)

    from oracle.torch_cpu_port import time_fn

    sec = time_fn(step, warmup, steps)
    return batch / sec, sec, cores


def cpu_index_rate(bq, warmup, steps, seed=2):
    import torch
    from oracle.torch_cpu_port import cpu_index_topk, time_fn

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(seed)
    corpus = torch.from_numpy((np.abs(rng.standard_normal((V_ARTICLES, JOINT))) * 0.1).astype(np.float32))
    q = torch.from_numpy(np.maximum(rng.standard_normal((bq, JOINT)) * 0.3, 0).astype(np.float32))
    sec = time_fn(lambda: cpu_index_topk(q, corpus, INDEX_K), warmup, steps)
    return bq / sec, sec, cores


def workload_name(batch):
    """config.workload: BASELINE.json configs[1], the same string on both arms."""
    cust = "1.37M" if V_CUSTOMERS == 1_371_980 else f"{V_CUSTOMERS / 1e6:g}M"
    return (f"c2: H&M-shaped two-tower ({cust} customers, 105k articles), id emb + joint dim {JOINT}, side features age/product_type/colour, "
            f"batch {batch}/GPU, logQ in-batch softmax, Adagrad lr 0.05")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    rate, sec, cores = cpu_train_rate(args.batch, warmup, steps)
    irate, isec, _ = cpu_index_rate(INDEX_BQ, 1, 3)
    sample = f"{steps} train steps of batch {args.batch} (same synthetic workload), torch-CPU port, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "examples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch),
                   "note": "reference TF is not installable here; this is the oracle restatement run with torch-CPU ops"},
        "cpu_baseline": {"value": rate, "unit": "examples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "index": {"value": irate, "unit": "queries/s", "ms_per_batch": isec * 1e3,
                  "config": f"N={V_ARTICLES}, E={JOINT}, K={INDEX_K}, Bq={INDEX_BQ}, torch-CPU matmul+topk, {cores} threads"},
    }
    print(json.dumps(line))



def index_stage_ms(lib, index, qe, reps=5):
    """Device time of the index search by stage (tt_debug_index_stages: CUDA events around each stage of a few extra, untimed calls)."""
    import ctypes

    import torch

    arr = (ctypes.c_float * 8)()
    torch.cuda.synchronize()
    lib.tt_debug_index_stages(arr)
    try:
        for _ in range(reps):
            index.search(qe)
        torch.cuda.synchronize()
    finally:
        lib.tt_debug_index_stages(None)
    names = ["prepare_queries", "threshold_pass", "select_threshold", "collect_pass", "rescore_sort", "exact_fallback"]
    return {n: arr[i] / reps for i, n in enumerate(names)}


# ---------------------------------------------------------------------------------------------------
# timing: blocks of exactly K steps, repeated until the timed region covers >= min_ms; the median block is reported
# ---------------------------------------------------------------------------------------------------
def _reduce(x, world, op="max"):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t)


def _gather_floats(x, world):
    if world == 1:
        return [x]
    import torch
    import torch.distributed as dist

    t = torch.zeros(world, device="cuda", dtype=torch.float64)
    t[dist.get_rank()] = x
    dist.all_reduce(t)
    return [round(float(v), 5) for v in t]


def _barrier(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def time_device_blocks(step, k, warmup, world, min_ms=1000.0, max_blocks=2000):
    """step(i) enqueues one step.  Returns the median per-step device time over blocks of k steps (max over ranks)."""
    import torch

    for i in range(warmup):
        step(i)
    _barrier(world)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(k):
        step(i)
    ev1.record()
    _barrier(world)
    probe = _reduce(ev0.elapsed_time(ev1), world)                       # sizes the region; every rank takes the same block count
    blocks = int(min(max_blocks, max(1, -(-min_ms // max(probe, 1e-3)))))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(blocks + 1)]
    _barrier(world)
    n = 0
    evs[0].record()
    for j in range(blocks):
        for i in range(k):
            step(n)
            n += 1
        evs[j + 1].record()
    _barrier(world)
    per = sorted(evs[j].elapsed_time(evs[j + 1]) for j in range(blocks))
    med_local = float(np.median(per))
    med = _reduce(med_local, world)
    region = _reduce(evs[0].elapsed_time(evs[blocks]), world)
    return {"ms_per_step": med / k, "blocks": blocks, "region_ms": region, "ms_per_step_mean": region / (blocks * k),
            "ms_per_step_min": _reduce(per[0], world) / k, "ms_per_step_max": _reduce(per[-1], world) / k,
            "ms_per_step_by_rank": _gather_floats(med_local / k, world)}


def time_host_blocks(step, k, warmup, world, min_ms=1000.0, max_blocks=2000):
    """step(i) runs one synchronous host-facing call (result on the host when it returns).  Wall clock, median block."""
    for i in range(warmup):
        step(i)
    _barrier(world)
    t0 = time.perf_counter()
    for i in range(k):
        step(i)
    probe = _reduce((time.perf_counter() - t0) * 1e3, world)
    blocks = int(min(max_blocks, max(1, -(-min_ms // max(probe, 1e-3)))))
    _barrier(world)
    per, n = [], 0
    t_start = time.perf_counter()
    for j in range(blocks):
        t0 = time.perf_counter()
        for i in range(k):
            step(n)
            n += 1
        per.append(time.perf_counter() - t0)
    region = time.perf_counter() - t_start
    _barrier(world)
    return {"sec_per_step": _reduce(float(np.median(per)), world) / k, "blocks": blocks, "region_ms": _reduce(region * 1e3, world)}

# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def build_gpu_model(attach=None):
    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory
    from pkg.schema import dtypes as tt
    from pkg.schema.features import Feature, FeatureFamily

    set_seed(1234)
    qf = [Feature("age", tt.float32, FeatureFamily.QUERY), Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=E_ID)]
    cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=E_ID),
          Feature("product_type_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=E_PTYPE),
          Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=E_COLOUR)]
    qf[1].set_vocab_size(V_CUSTOMERS); cf[0].set_vocab_size(V_ARTICLES); cf[1].set_vocab_size(V_PTYPE); cf[2].set_vocab_size(V_COLOUR)
    probs = article_probs()
    lookup = {str(i + 1): float(p) for i, p in enumerate(probs)}
    model = TwoTowerModel(qf, cf, "article_id", JOINT, candidate_prob_lookup=lookup)
    if attach is not None:      # data parallel: shard the (still unmaterialised) tables first, so the owners initialise their shards
        attach(model)           # and the optimizer slots are only ever allocated shard-sized
    model.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
    return model


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pkg import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = N.load()
    pk = peaks()
    B, K, W = args.batch, max(args.steps, 1), max(args.warmup, 3)
    attach = None
    if world > 1:
        from pkg.modelling.distributed import DataParallel

        attach = lambda m: DataParallel(m, shard_tables=(args.tables == "sharded"), global_negatives=args.global_negatives)   # noqa: E731
    torch.cuda.reset_peak_memory_stats()
    model = build_gpu_model(attach)
    if args.simt:
        model.impl = N.TT_IMPL_SIMT
    model.use_cuda_graph = not args.no_graph   # data parallel: two captured compute phases around the eager NCCL exchanges
    rng = np.random.default_rng(1000 + rank)
    pool = 8
    host_batches = [make_batch(rng, B) for _ in range(pool)]
    dev_batches = [{k: torch.from_numpy(v).cuda() for k, v in hb.items()} for hb in host_batches]
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in hb.items()} for hb in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches per step (eager, before any graph capture decision matters)
    c0 = lib.tt_launch_count()
    model.train_step(dev_batches[0])
    torch.cuda.synchronize()
    launches_per_step = int(lib.tt_launch_count() - c0)

    with ClockSampler(local) as clk:
        dev = time_device_blocks(lambda i: model.train_step(dev_batches[i % pool]), K, W, world, args.min_ms)
    ms_per_step = dev["ms_per_step"]
    value = world * B / (ms_per_step * 1e-3)
    clocks = clk.summary()

    # end-to-end through the public API: pinned host ids -> H2D -> step -> loss D2H, every step
    e2e_step, last = pipelined_loss_step(model, pinned)
    host = time_host_blocks(e2e_step, K, 3, world, args.min_ms)
    loss = last[0]
    e2e = world * B / host["sec_per_step"]

    line = {
        "metric": METRIC, "value": value, "unit": "examples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": ("tf32 tower products; softmax on fp16 operand tiles (per-tensor power-of-two scaling), fp32 accumulate, fp32 positive term"
                                                     if model._tc_ok() and model.impl != N.TT_IMPL_SIMT else "f32"),
        "data": "synthetic",
        "config": {"workload": workload_name(B),
                   "parallelism": f"dp{world}" if world > 1 else "single", "cuda_graph": bool(model.use_cuda_graph),
                   "l2": "inputs larger than L2: tables + Adagrad accumulators 0.76 GB, random rows each step; no explicit flush",
                   "last_loss": loss},
        "e2e": {"value": e2e, "unit": "examples/s", "h2d_bytes_per_step": B * BYTES_PER_EXAMPLE_H2D, "d2h_bytes_per_step": 4,
                "blocks": host["blocks"], "timed_region_ms": host["region_ms"],
                "how": "model.train_step on pinned host columns (staged by one kernel reading them in place), loss of every step copied to "
                       "pinned memory and read on the host one step late"},
        "gpu_launches": launches_per_step * K, "clocks": clocks,
        "timing": {"blocks": dev["blocks"], "steps_per_block": K, "ms_per_step_by_rank": dev["ms_per_step_by_rank"], "timed_region_ms": dev["region_ms"], "ms_per_step_mean": dev["ms_per_step_mean"],
                   "ms_per_step_min_block": dev["ms_per_step_min"], "ms_per_step_max_block": dev["ms_per_step_max"],
                   "note": "ms_per_step = median over back-to-back blocks of exactly `steps` steps (CUDA events on the launching stream, "
                           "max over ranks); the blocks together cover >= --min-ms of device time"},
    }

    line["config"]["hbm_peak_gb_per_gpu"] = round(torch.cuda.max_memory_allocated() / 1e9, 2)
    if world > 1:
        line["config"]["tables"] = "row-sharded over the GPUs (rows read / gradient rows pulled over NVLink peer memory)" if model.dist.shard_tables \
            else "replicated (all-gathered gradient rows)"
        if args.global_negatives:
            line["config"]["negatives"] = f"cross-GPU: {world * B} candidate columns per query row (all-gathered over NVLink peer memory)"
        barrier()       # every rank's last table update has landed before anybody embeds the corpus
    line["dp_phases_ms" if world > 1 else "step_phases_ms"] = dp_phase_times(model, B, dev_batches)
    barrier()
    if rank == 0:
        line["roofline"] = softmax_roofline(model, B, pk, lib, world)
        if not args.no_hbm:
            line["hbm_kernels"] = hbm_rooflines(pk, lib)
    line["index"] = index_bench(model, pk, lib, K, world, args)   # every rank takes part (row-sharded corpus when N > 1)
    if not args.no_big_index:
        line["index"]["row_sharded_large"] = big_index_legs(model, pk, lib, K, world, rank, args)
    if not args.no_c3 and not args.global_negatives and B != C3_BATCH:
        # after the index legs: a few hundred more SUM-loss steps at 8x the batch, unscheduled lr, collapse this synthetic model onto
        # the Zipf head, and an index over a collapsed corpus measures the exact-fallback path instead of the filter
        line["c3"] = c3_leg(model, args, pk, lib, world, rank)     # BASELINE configs[2]: the same model at batch 65536 per GPU
    if world > 1 and not args.no_parity:
        line["parity"] = multi_gpu_parity(world, rank, args)
        line["parity_ok"] = bool(all(v.get("ok") for v in line["parity"].values())
                                 and all(l.get("parity_ok", True) for l in line["index"].get("row_sharded_large", [])))
    if world == 1:
        if not args.no_cpu:
            rate, sec, cores = cpu_train_rate(B, 2, args.cpu_steps)
            line["cpu_baseline"] = {"value": rate, "unit": "examples/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_steps} train steps of batch {B} of the same workload (torch-CPU oracle port, {cores} threads, {sec * 1e3:.1f} ms/step)"}
            irate, isec, _ = cpu_index_rate(INDEX_BQ, 1, 3)
            line["index"]["cpu_baseline"] = {"value": irate, "unit": "queries/s", "cores": cores, "kind": "port",
                                             "sample": f"3 batches of {INDEX_BQ} queries, torch-CPU matmul+topk ({isec * 1e3:.0f} ms/batch)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



C3_BATCH = 65536


def pipelined_loss_step(model, pinned):
    """The end-to-end step: pinned host columns in (read in place by the one staging kernel), train_step, and the step's loss copied
    to pinned host memory and READ on the host -- one step late, while the next step already runs (two slots, one event each), so
    the read does not drain the GPU.  Every step's loss is read exactly once."""
    import torch

    slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    pending = [False, False]
    last = [0.0]
    pool = len(pinned)

    def step(i):
        s = i & 1
        out = model.train_step(pinned[i % pool])
        slots[s].copy_(out["loss"].reshape(1), non_blocking=True)
        evs[s].record()
        pending[s] = True
        if pending[1 - s]:
            evs[1 - s].synchronize()
            last[0] = float(slots[1 - s][0])
            pending[1 - s] = False

    return step, last


def c3_leg(model, args, pk, lib, world, rank):
    """BASELINE configs[2] ("c3"): the same towers at batch 65536 per GPU -- 64x the logits of c2 per step, the shape where the
    in-batch softmax dominates.  Same timing rules as the headline leg."""
    import torch

    B = C3_BATCH
    rng = np.random.default_rng(4000 + rank)
    pool = 4
    host = [make_batch(rng, B) for _ in range(pool)]
    devb = [{k: torch.from_numpy(v).cuda() for k, v in hb.items()} for hb in host]
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in hb.items()} for hb in host]
    k = max(2, min(args.steps, 10))
    dev = time_device_blocks(lambda i: model.train_step(devb[i % pool]), k, 3, world, args.min_ms / 2)
    e2e_step, last = pipelined_loss_step(model, pinned)
    hostt = time_host_blocks(e2e_step, k, 2, world, args.min_ms / 2)
    out = {"workload": workload_name(B).replace("c2:", "c3:"), "value": world * B / (dev["ms_per_step"] * 1e-3), "unit": "examples/s",
           "ms_per_step": dev["ms_per_step"], "steps": k, "blocks": dev["blocks"], "timed_region_ms": dev["region_ms"],
           "e2e": {"value": world * B / hostt["sec_per_step"], "unit": "examples/s", "h2d_bytes_per_step": B * BYTES_PER_EXAMPLE_H2D,
                   "d2h_bytes_per_step": 4}, "last_loss": last[0]}
    if rank == 0:
        out["roofline"] = softmax_roofline(model, B, pk, lib)
    return out


# ---------------------------------------------------------------------------------------------------
# row-sharded index at 1e7 / 1e8 rows (BASELINE configs[3]): every shard is generated ON its owner
# ---------------------------------------------------------------------------------------------------
CORPUS_CHUNK = 1 << 20


def corpus_rows(lo, hi, e=JOINT):
    """Rows [lo, hi) of the synthetic corpus, generated on this GPU.  Row values depend only on the global row number (chunks of
    2^20 rows, one seeded generator per chunk), so any sharding of the same corpus holds the same rows."""
    import torch

    out = torch.empty((hi - lo, e), dtype=torch.float32, device="cuda")
    c = lo // CORPUS_CHUNK
    while c * CORPUS_CHUNK < hi:
        a, b = c * CORPUS_CHUNK, (c + 1) * CORPUS_CHUNK
        g = torch.Generator(device="cuda").manual_seed(90_000 + c)
        chunk = torch.randn((CORPUS_CHUNK, e), generator=g, device="cuda") * 0.25
        s0, s1 = max(a, lo), min(b, hi)
        out[s0 - lo:s1 - lo].copy_(chunk[s0 - a:s1 - a])
        del chunk
        c += 1
    return out


def big_index_legs(model, pk, lib, steps, world, rank, args):
    import torch

    from pkg.modelling.distributed import shard_bounds
    from pkg.modelling.indices.brute_force import BruteForceIndex

    if args.big_index_rows == "auto":
        sizes = [10_000_000] + ([100_000_000] if world >= 4 else [])
    else:
        sizes = [int(float(x)) for x in args.big_index_rows.split(",") if x]
    rng = np.random.default_rng(177)                      # the SAME queries on every rank (row-sharded corpus: queries are replicated)
    pool = 4
    hq = [{"age": rng.random((INDEX_BQ, 1)).astype(np.float32), "customer_id": rng.integers(1, V_CUSTOMERS + 1, size=(INDEX_BQ, 1)).astype(np.int32)}
          for _ in range(pool)]
    dq = [{k: torch.from_numpy(v).cuda() for k, v in h.items()} for h in hq]
    pq = [{k: torch.from_numpy(v).pin_memory() for k, v in h.items()} for h in hq]
    res = []
    for n_rows in sizes:
        lo, hi = shard_bounds(n_rows, rank, world)
        t0 = time.perf_counter()
        rows = corpus_rows(lo, hi)
        index = BruteForceIndex.from_local_rows(INDEX_K, model.query_tower, rows, lo, n_rows, identifiers=None, world=world)
        index.impl = model.impl
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        k = max(2, min(steps, 10))
        dev = time_device_blocks(lambda i: index.query_indices(dq[i % pool]), k, 2, world, args.min_ms / 2)
        qe = index._embed_queries(dq[0])
        kern = time_device_blocks(lambda i: index.search(qe), k, 1, world, args.min_ms / 2)
        outs = [torch.empty((INDEX_BQ, INDEX_K), dtype=torch.int32).pin_memory() for _ in range(2)]
        pend = [None, None]

        def call(i, index=index, pend=pend):
            s = i & 1
            pend[s] = index(pq[i % pool], out=outs[s], wait=False)
            if pend[1 - s] is not None:
                pend[1 - s][1].synchronize()
                pend[1 - s] = None

        host = time_host_blocks(call, k, 2, world, args.min_ms / 2)
        torch.cuda.synchronize()
        flops = 2.0 * INDEX_BQ * n_rows * JOINT / world
        ach = flops / (kern["ms_per_step"] * 1e-3) / 1e12
        leg = {"rows": n_rows, "value": INDEX_BQ / (dev["ms_per_step"] * 1e-3), "unit": "queries/s", "ms_per_batch": dev["ms_per_step"],
               "search_only_ms": kern["ms_per_step"], "e2e": {"value": INDEX_BQ / host["sec_per_step"], "unit": "queries/s",
                                                              "h2d_bytes_per_step": INDEX_BQ * 8, "d2h_bytes_per_step": INDEX_BQ * INDEX_K * 4},
               "blocks": dev["blocks"], "timed_region_ms": dev["region_ms"], "shard_build_s": _reduce(build_s, world),
               "stage_ms": index_stage_ms(lib, index, qe, 2),
               "config": f"{n_rows} x {JOINT} fp32 synthetic corpus ({n_rows * JOINT * 4 / 1e9:.1f} GB + prepared copy), K={INDEX_K}, Bq={INDEX_BQ} replicated queries; "
                         + (f"rows sharded over {world} GPUs, each shard generated on its owner (no rank holds the corpus), per-shard top-K all-gathered "
                            f"(NCCL) and merged on the device" if world > 1 else "single shard"),
               "roofline": {"bound": "tensor", "kernel": "index scoring + top-K (per GPU, incl. the merge when sharded)", "achieved": ach,
                            "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": ach / pk["tflops_burst"], "traffic": None,
                            "hbm_gbs_corpus_read": (hi - lo) * JOINT * 4 * ((INDEX_BQ + 127) // 128) / (kern["ms_per_step"] * 1e-3) / 1e9,
                            "algorithmic_flop": flops}}
        if world > 1 and not args.no_parity:
            # merged top-K of the sharded corpus == top-K of the unsharded corpus, held by rank 0 alone for this check
            s_sh, i_sh = index.search(qe)
            ok = 1.0
            if rank == 0 and n_rows * JOINT * 8 < 60e9:
                whole = BruteForceIndex.from_local_rows(INDEX_K, model.query_tower, corpus_rows(0, n_rows), 0, n_rows)
                whole.impl = model.impl
                s_1, i_1 = whole.search(qe)
                ok = float(bool(torch.equal(i_1, i_sh)) and bool(torch.equal(s_1, s_sh)))
                del whole
            leg["parity_ok"] = bool(-_reduce(-ok, world) >= 1.0)
            leg["parity_check"] = "merged (scores, row indices) of the sharded search bit-equal to the unsharded search of the same corpus on rank 0"
        res.append(leg)
        del index, rows, qe
        torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------------
# in-process multi-GPU parity (world > 1): the N-rank step against single-GPU steps of the same library
# ---------------------------------------------------------------------------------------------------
def _small_model(seed, lr, acc0):
    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory
    from pkg.schema import dtypes as tt
    from pkg.schema.features import Feature, FeatureFamily

    set_seed(seed)
    qf = [Feature("age", tt.float32, FeatureFamily.QUERY), Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=E_ID)]
    cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=E_ID),
          Feature("product_type_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=E_PTYPE),
          Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=E_COLOUR)]
    qf[1].set_vocab_size(PAR_CUSTOMERS); cf[0].set_vocab_size(PAR_ARTICLES); cf[1].set_vocab_size(V_PTYPE); cf[2].set_vocab_size(V_COLOUR)
    lookup = {str(i + 1): 1.0 / PAR_ARTICLES for i in range(PAR_ARTICLES)}
    model = TwoTowerModel(qf, cf, "article_id", JOINT, candidate_prob_lookup=lookup)
    model.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": lr, "initial_accumulator_value": acc0}))
    return model


PAR_CUSTOMERS, PAR_ARTICLES, PAR_BATCH = 50_000, 20_000, 2048


def _par_batch(r, b=PAR_BATCH):
    rng = np.random.default_rng(31_000 + r)
    art = rng.integers(1, PAR_ARTICLES + 1, size=b).astype(np.int32)
    return {"age": rng.random((b, 1)).astype(np.float32), "customer_id": rng.integers(1, PAR_CUSTOMERS + 1, size=(b, 1)).astype(np.int32),
            "article_id": art.reshape(b, 1), "product_type_name": (art % V_PTYPE + 1).astype(np.int32).reshape(b, 1),
            "colour_group_name": (art % V_COLOUR + 1).astype(np.int32).reshape(b, 1)}


def _flat_state(model):
    """[dense parameters | every table, full rows] as one fp32 device vector (collective when the tables are row-sharded)."""
    import torch

    parts = [model._store.params[: int(model._store.used)].detach().reshape(-1)]
    seen = set()
    for _, _, t in model._tables():
        if id(t) not in seen:
            seen.add(id(t))
            parts.append(t.full_weight().detach().reshape(-1))
    return torch.cat(parts).clone()


def multi_gpu_parity(world, rank, args):
    """Two checks, both against single-GPU runs of this same library (whose parity with the oracle tests/ establishes):
      dp_step       -- Adagrad with a 1e12 accumulator is a fixed-rate SGD to fp32 rounding, so the data-parallel update must equal
                       the SUM over ranks of the updates single-GPU replicas make on each rank's batch from the same weights,
                       and the summed loss the sum of their losses.
      global_negatives -- the N-rank step with cross-GPU negatives is the single-GPU step on the concatenated batch."""
    import torch
    import torch.distributed as dist

    from pkg.modelling.distributed import DataParallel

    out = {}
    lr, acc0 = 0.05 * 1e6, 1e12
    try:
        ref = _small_model(77, lr, acc0)
        w0 = _flat_state(ref)
        mine = {k: torch.from_numpy(v).cuda() for k, v in _par_batch(rank).items()}
        loss_r = float(ref.train_step(mine)["loss"])
        delta = _flat_state(ref) - w0
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
        loss_sum = _reduce(loss_r, world, "sum")
        dp = _small_model(77, lr, acc0)
        DataParallel(dp, shard_tables=(args.tables == "sharded"))
        loss_dp = _reduce(float(dp.train_step(mine)["loss"]), world, "sum")
        dp.dist.barrier()
        got = _flat_state(dp) - w0
        err = float((got - delta).abs().max() / delta.abs().max().clamp_min(1e-30))
        err = _reduce(err, world)
        lerr = abs(loss_dp - loss_sum) / abs(loss_sum)
        out["dp_step"] = {"ok": bool(err <= 2e-4 and lerr <= 1e-5), "max_update_err_rel_to_max_update": err, "loss_rel_err": lerr,
                          "loss": loss_dp, "check": f"{world}-rank step (batch {PAR_BATCH}/rank, {args.tables} tables) == sum of {world} single-GPU replica "
                                                    "updates from the same weights (linearised Adagrad), all parameters and every table row"}
        del ref, dp
    except Exception as ex:      # a failed check must not lose the timing line
        out["dp_step"] = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}
    # cross-GPU negatives, twice: on the exact fp32 path (tight: proves the plumbing -- candidate all-gather, diagonal offsets,
    # reduce-scatter of dC, sharded updates; 1e-4: fp32 sums over 16384 rows in two different orders) and on the tensor-core path
    # (the product path; its fp16 operand tiles are scaled per call, so the N-rank and the single-GPU evaluation round differently:
    # tables to 2e-3 of the largest update; the dense GRADIENT is a batch-wide sum of up to 16384 independently rounded rows that
    # largely cancel -- measured 2.5e-3 of its norm between the two evaluations at 8 ranks -- and gets 5e-3)
    from pkg import _native as N

    for name, impl, tol_tab, tol_grad in (("global_negatives_exact", N.TT_IMPL_SIMT, 1e-4, 1e-4), ("global_negatives", N.TT_IMPL_AUTO, 2e-3, 5e-3)):
        try:
            if args.tables != "sharded":
                break
            whole = {k: torch.from_numpy(np.concatenate([_par_batch(r)[k] for r in range(world)], axis=0)).cuda() for k in _par_batch(0)}
            ref = _small_model(78, 0.05, 0.1)
            ref.impl = impl
            nd = int(ref._store.used)
            w0 = _flat_state(ref)
            loss_1 = float(ref.train_step(whole)["loss"])
            want = _flat_state(ref)
            g_ref = ref._store.grads[:nd].clone()
            gn = _small_model(78, 0.05, 0.1)
            gn.impl = impl
            DataParallel(gn, shard_tables=True, global_negatives=True)
            mine = {k: torch.from_numpy(v).cuda() for k, v in _par_batch(rank).items()}
            loss_g = _reduce(float(gn.train_step(mine)["loss"]), world, "sum")
            gn.dist.barrier()
            got = _flat_state(gn)
            g_dp = gn._store.grads[:nd].clone()
            upd = (want - w0)[nd:].abs().max().clamp_min(1e-30)
            err_tab = _reduce(float((got - want)[nd:].abs().max() / upd), world)
            err_dense_upd = _reduce(float((got - want)[:nd].abs().max() / (want - w0)[:nd].abs().max().clamp_min(1e-30)), world)
            err_grad = _reduce(float((g_dp - g_ref).norm() / g_ref.norm().clamp_min(1e-30)), world)
            lerr = abs(loss_g - loss_1) / abs(loss_1)
            out[name] = {"ok": bool(err_tab <= tol_tab and err_grad <= tol_grad and lerr <= 1e-5), "table_update_err_rel_to_max_update": err_tab,
                         "dense_gradient_l2_rel_err": err_grad, "dense_update_err_rel_to_max_update": err_dense_upd, "loss_rel_err": lerr,
                         "loss": loss_g, "path": "exact fp32 (CUDA cores)" if impl == N.TT_IMPL_SIMT else "tensor cores (fp16 operand tiles)",
                         "tolerances": {"tables": tol_tab, "dense_gradient": tol_grad, "loss": 1e-5},
                         "check": f"{world}-rank step with cross-GPU negatives ({world * PAR_BATCH} columns per row) == the single-GPU step on the "
                                  f"concatenated batch of {world * PAR_BATCH} (Adagrad lr 0.05): every table row, the summed dense gradient, the loss"}
            del ref, gn
        except Exception as ex:
            out[name] = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}
    torch.cuda.empty_cache()
    return out


def dp_phase_times(model, B, dev_batches, n=200):
    """Device time of the four phases of a data-parallel step measured INSIDE the real step graph: the step is re-captured with
    tt_stamp kernels (%globaltimer) between the phases and n steps run back to back, like the timed region.  sync_* is what the
    device barrier costs this rank, including waiting for the slowest one."""
    import torch

    ring_len = n
    ring = torch.zeros(1 + ring_len * 5, dtype=torch.int64, device="cuda")
    model.phase_stamps = ring
    saved = model._steps.pop(B, None)      # re-capture this batch shape with the stamps in the graph
    try:
        for i in range(4):                 # eager step, capture, two replays
            model.train_step(dev_batches[i % len(dev_batches)])
        torch.cuda.synchronize()
        ring.zero_()
        for i in range(n):
            model.train_step(dev_batches[i % len(dev_batches)])
        torch.cuda.synchronize()
        t = ring[1:].view(ring_len, 5).double().cpu().numpy()
    finally:
        model.phase_stamps = None
        model._steps.pop(B, None)
        if saved is not None:
            model._steps[B] = saved
    d = np.diff(t, axis=1) * 1e-6          # ms
    gap = (t[1:, 0] - t[:-1, 4]) * 1e-6   # end of a step -> start of the next (graph launch gap)
    med = np.median(d, axis=0)
    return {"sync_ids": float(med[0]), "phase_a": float(med[1]), "sync_grads": float(med[2]), "phase_b": float(med[3]),
            "between_steps": float(np.median(gap)), "steps": n,
            "note": "this rank, medians over back-to-back steps, stamps inside the captured step graph (each stamp kernel adds ~2 us); "
                    "sync_* = device barrier incl. waiting for the slowest rank"}


def softmax_roofline(model, B, pk, lib, world=1):
    """Dominant kernel group: in-batch softmax fwd + bwd (6.B.Bc.E algorithmic flop; Bc = B, or world.B candidate columns with cross-GPU
    negatives), timed alone with CUDA events."""
    import torch

    from pkg import _native as N

    sw = model._step_ws(B)
    e = model.joint_embedding_size
    use_tc = model.impl != N.TT_IMPL_SIMT and model._tc_ok()
    q, c = (sw.q.out_tf32, sw.c.out_tf32) if use_tc else (sw.q.acts[-1], sw.c.acts[-1])
    impl = N.TT_IMPL_TC if use_tc else N.TT_IMPL_SIMT
    bias = sw.col_bias.data_ptr() if sw.col_bias is not None else None
    st = N.stream_ptr()
    gn = model.dist is not None and model.dist.global_negatives
    bc = world * B if gn else B
    dc, off = sw.dc, 0
    if gn:      # the shape the step runs: this rank's B rows against the world.B candidates of all ranks (synthetic candidates here)
        g = torch.Generator(device="cuda").manual_seed(11)
        c = torch.relu(torch.randn((bc, e), generator=g, device="cuda") * 0.3)
        bias_t = torch.log(torch.rand(bc, generator=g, device="cuda") * 0.01 + 1e-5)
        bias = bias_t.data_ptr()
        dc = torch.empty((bc, e), dtype=torch.float32, device="cuda")
    ws = torch.empty(int(lib.tt_softmax_workspace_bytes(B, bc, e)), dtype=torch.uint8, device="cuda")

    def once():   # the call the train step makes: operand prep, pass 1 (forward + dQ), combine, pass 2 (dC), combine
        N.check(lib.tt_inbatch_softmax_step(q.data_ptr(), e, c.data_ptr(), e, bias, B, bc, e, off, sw.lse.data_ptr(), sw.loss.data_ptr(),
                                            sw.dq.data_ptr(), e, dc.data_ptr(), e, ws.data_ptr(), ws.numel(), impl, st))

    t = time_device_blocks(lambda i: once(), 10, 3, 1, 300.0)      # rank 0 alone: no collective in the timing helper
    sec = t["ms_per_step"] * 1e-3
    flops = 6.0 * B * bc * e
    achieved = flops / sec / 1e12
    return {"bound": "tensor", "kernel": "in-batch softmax fwd+bwd, %d x %d logits (%s)" % (B, bc, "tcgen05, fp16 operand tiles" if use_tc else "fp32 CUDA cores"),
            "achieved": achieved, "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_burst"],
            "traffic": NCU_SOFTMAX_DRAM_BYTES if (use_tc and B == 8192 and e == 64 and NCU_SOFTMAX_DRAM_BYTES) else None, "traffic_source": NCU_SOFTMAX_SOURCE,
            "ms": sec * 1e3, "algorithmic_flop": flops, "peak_source": pk["source"] + ", dense bf16 burst (kernel timed alone)"}


# dram__bytes_read.sum + dram__bytes_write.sum of the six launches of one softmax call (amax 4.20, convert 4.24, pass 1 2.79 + 0.06,
# combine1 25.95, pass 2 2.21, combine2 25.21 MB) from the round-2 `ncu --set full` capture at B = 8192, E = 64.  ncu replays every
# kernel cold: the two combine kernels' ~25 MB are the per-CTA partial accumulators the passes just wrote, L2 hits in the running
# step.  Algorithmic bytes: 4 MB of fp32 operands in, 4 MB of gradients out.  A constant from that capture, NOT measured in this run.
NCU_SOFTMAX_DRAM_BYTES = 4204032 + 4235776 + 2787840 + 62976 + 25946368 + 2207744 + 256 + 25209600
NCU_SOFTMAX_SOURCE = ("profiles/r02_flash_b8192_ncu_full.md (ncu --set full, one call = 6 launches, cold-cache replay; constant from that capture, "
                      "not measured in this run)")


# dram__bytes_read.sum + dram__bytes_write.sum per launch at the default size of hbm_rooflines (2^20 ids, 1.37 M x 64 table):
# gather 245.1 + 222.6 MB (duplicate ids hit L2, part of the output is still in L2 when the kernel ends); segmented reduce
# 655.0 + 344.9 MB plus the combine kernel's 14.4 MB, for 1.02 GB of algorithmic bytes
NCU_HBM_DRAM_BYTES = {"gather": 245108480 + 222646272, "update": 655021568 + 344898048 + 14365696}
NCU_HBM_SOURCE = ("profiles/r01d_hbm_kernels_ncu_full.md (gather), profiles/r01e_sparse_block_ncu_full.md (update); ncu --set full, per launch; "
                  "constants from those captures, not measured in this run")

HBM_TIMING = None     # (warm-up launches, timed launches) override used by scripts/hbm_microbench.py --once under ncu


def hbm_rooflines(pk, lib, b=1 << 20, e=64, rows=V_CUSTOMERS + 1):
    """The HBM-bound kernels at a batch large enough to leave the launch-latency regime (2^20 ids; the step's own batch moves
    only 5 MB): embedding gather (4.e B read + 4.e B written + 4 B id per example) and the de-duplicated sparse Adagrad
    ((1 + 4.U/B).4.e B per example, U = unique rows).  Timed alone with CUDA events."""
    import torch

    from pkg import _native as N
    from pkg.modelling._device import feature_array

    g = torch.Generator(device="cuda").manual_seed(5)
    table = torch.rand((rows, e), generator=g, device="cuda") * 0.1 - 0.05
    acc = torch.full_like(table, 0.1)
    ids = torch.randint(1, rows, (b,), generator=g, device="cuda", dtype=torch.int32)
    grad = torch.randn((b, e), generator=g, device="cuda") * 1e-3
    out = torch.empty((b, e), dtype=torch.float32, device="cuda")
    uniq = int(torch.unique(ids).numel())
    st = N.stream_ptr()
    feats = feature_array([dict(table=table.data_ptr(), src=ids.data_ptr(), rows=rows, e=e, col=0)])
    jobs = (N.TTSparseJob * 1)()
    jobs[0].table, jobs[0].slot0, jobs[0].slot1 = table.data_ptr(), acc.data_ptr(), None
    jobs[0].rows, jobs[0].e, jobs[0].nsrc, jobs[0].n_per_src = rows, e, 1, b
    jobs[0].ids[0], jobs[0].grad[0], jobs[0].grad_ld[0] = ids.data_ptr(), grad.data_ptr(), e
    ws = torch.empty(int(lib.tt_sparse_workspace_bytes(1, b, e)), dtype=torch.uint8, device="cuda")

    def gather():
        N.check(lib.tt_gather_concat(feats, 1, b, e, out.data_ptr(), e, st), "tt_gather_concat")

    def sort():
        N.check(lib.tt_sparse_sort(jobs, 1, ws.data_ptr(), ws.numel(), st), "tt_sparse_sort")

    def update():
        N.check(lib.tt_sparse_adagrad(jobs, 1, 0.05, 1e-7, ws.data_ptr(), ws.numel(), st), "tt_sparse_adagrad")

    def timeit(fn, n=10, pre=None):
        warm, n = HBM_TIMING if HBM_TIMING else (3, n)
        for _ in range(warm):
            if pre:
                pre()
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n):
            if pre:
                pre()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            tot += ev0.elapsed_time(ev1)
        return tot * 1e-3 / n

    res = []
    default_size = (b, e, rows) == (1 << 20, 64, V_CUSTOMERS + 1)     # the size the committed ncu captures were taken at
    for name, fn, pre, nbytes, cap in (
            ("embedding gather (tt_gather_concat)", gather, None, b * (8.0 * e + 4), "gather"),
            ("sparse Adagrad: id sort + segmented reduce + row update", lambda: (sort(), update()), None, b * (1 + 4.0 * uniq / b) * 4 * e, None),
            ("sparse Adagrad: segmented reduce + row update (ids already sorted)", update, sort, b * (1 + 4.0 * uniq / b) * 4 * e, "update")):
        sec = timeit(fn, pre=pre)
        ach = nbytes / sec / 1e9
        res.append({"bound": "hbm", "kernel": name, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": NCU_HBM_DRAM_BYTES[cap] if (cap and default_size) else None,
                    "traffic_source": NCU_HBM_SOURCE if (cap and default_size) else None, "ms": sec * 1e3, "algorithmic_bytes": nbytes,
                    "config": f"{b} ids into a {rows} x {e} fp32 table ({uniq} unique rows); table + accumulator 0.7 GB > L2"})
    return res


def index_bench(model, pk, lib, steps, world=1, args=None):
    """Index half of the metric: N=105 542 candidate-tower outputs, E=64, top-100, 2048 queries per batch.
    N > 1 ranks, two sharding modes:
      * headline `value`: the 27 MB corpus is replicated and the QUERIES are sharded (every rank answers its own
        2048-query batches; no exchange on the data path) -- the natural layout when the corpus fits one GPU;
      * `row_sharded`: the corpus is split row-wise, every rank scores all queries against its shard, the per-shard
        top-K lists are all-gathered over NCCL and merged on the device (the layout the 10M / 100M-row configs need)."""
    import torch
    import torch.distributed as dist

    from pkg import _native as N
    from pkg.modelling.indices.brute_force import BruteForceIndex
    from pkg.modelling.distributed import make_sharded_index

    art = np.arange(1, V_ARTICLES + 1, dtype=np.int32)
    pairs = []
    for lo in range(0, V_ARTICLES, 10000):   # candidate_batch_size = 10000 (training_config.py:36)
        a = art[lo:lo + 10000]
        x = {"article_id": a.reshape(-1, 1), "product_type_name": (a % V_PTYPE + 1).reshape(-1, 1), "colour_group_name": (a % V_COLOUR + 1).reshape(-1, 1)}
        pairs.append((a, model.candidate_tower(x)))
    rank = int(os.environ.get("RANK", "0"))
    rng = np.random.default_rng(77 + rank)
    pool = 4
    hq = [{"age": rng.random((INDEX_BQ, 1)).astype(np.float32), "customer_id": rng.integers(1, V_CUSTOMERS + 1, size=(INDEX_BQ, 1)).astype(np.int32)}
          for _ in range(pool)]
    dq = [{k: torch.from_numpy(v).cuda() for k, v in h.items()} for h in hq]
    pq = [{k: torch.from_numpy(v).pin_memory() for k, v in h.items()} for h in hq]
    n = max(steps, 5)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    min_ms = (args.min_ms if args is not None else 1000.0) / 2
    outs = [torch.empty((INDEX_BQ, INDEX_K), dtype=torch.int32).pin_memory() for _ in range(2)]

    def measure(index):
        c0 = lib.tt_launch_count()
        for i in range(3):
            index.query_indices(dq[i % pool])
        torch.cuda.synchronize()
        per_call = int((lib.tt_launch_count() - c0) // 3)
        dev = time_device_blocks(lambda i: index.query_indices(dq[i % pool]), n, 2, world, min_ms)
        # scoring + selection alone (embeddings resident): the dominant kernel group of the index path
        qe = index._embed_queries(dq[0])
        kern = time_device_blocks(lambda i: index.search(qe), n, 2, world, min_ms)
        # host-facing call: pinned host ids in, (Bq, K) identifiers out in a caller-owned pinned buffer
        ids = [None]
        pending = [None, None]

        def call(i):          # submit batch i, then read batch i-1 (its D2H overlaps this batch's kernels); every result is read
            s = i & 1
            pending[s] = index(pq[i % pool], out=outs[s], wait=False)
            if pending[1 - s] is not None:
                view, done = pending[1 - s]
                done.synchronize()
                ids[0] = view
                pending[1 - s] = None

        host = time_host_blocks(call, n, 3, world, min_ms)
        torch.cuda.synchronize()
        stages[0] = index_stage_ms(lib, index, qe)
        by_rank[0] = {"query_tower_and_search_ms": dev["ms_per_step_by_rank"], "search_only_ms": kern["ms_per_step_by_rank"]}
        return dev["ms_per_step"] * 1e-3, kern["ms_per_step"] * 1e-3, host["sec_per_step"], per_call, ids[0].copy()

    stages = [None]
    by_rank = [None]

    replicated = BruteForceIndex(INDEX_K, model.query_tower, pairs)
    replicated.impl = model.impl
    sec, ksec, e2e_sec, per_call, ids = measure(replicated)
    flops = 2.0 * INDEX_BQ * V_ARTICLES * JOINT
    ach = flops / ksec / 1e12
    out = {"metric": "index queries/s (top-100, 105k items)", "value": world * INDEX_BQ / sec, "unit": "queries/s", "ms_per_batch": sec * 1e3,
           "e2e": {"value": world * INDEX_BQ / e2e_sec, "unit": "queries/s", "h2d_bytes_per_step": INDEX_BQ * 8, "d2h_bytes_per_step": INDEX_BQ * INDEX_K * 4,
                   "how": "index(queries, out=pinned, wait=False) on pinned host columns; batch k's identifiers are read on the host after batch "
                          "k+1 has been submitted (two result buffers, one event each)"},
           "config": f"N={V_ARTICLES} candidate-tower rows, E={JOINT}, K={INDEX_K}, Bq={INDEX_BQ} per GPU; corpus 27 MB is L2-resident (stated); "
                     + (f"corpus replicated, queries sharded over {world} GPUs (no data-path collective)" if world > 1 else "single shard"),
           "gpu_launches_per_batch": per_call,
           "roofline": {"bound": "tensor", "kernel": "index scoring + top-K (per GPU)", "achieved": ach, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                        "frac": ach / pk["tflops_burst"], "traffic": None, "ms": ksec * 1e3, "algorithmic_flop": flops},
           "stage_ms": stages[0], "by_rank": by_rank[0], "sample_ids": [str(x) for x in ids[0, :3]]}
    if world > 1:
        sharded = make_sharded_index(INDEX_K, model.query_tower, pairs)
        sharded.impl = model.impl
        ssec, sksec, se2e, sper, _ = measure(sharded)
        out["row_sharded"] = {"value": INDEX_BQ / ssec, "unit": "queries/s", "ms_per_batch": ssec * 1e3, "e2e": INDEX_BQ / se2e,
                              "config": f"corpus row-sharded over {world} GPUs, every rank scores all {INDEX_BQ} queries, NCCL all-gather + on-device merge",
                              "gpu_launches_per_batch": sper}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--simt", action="store_true", help="force the exact fp32 CUDA-core contraction path")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--global-negatives", action="store_true",
                    help="--gpus > 1: every rank scores its batch against the candidates of ALL ranks (BASELINE configs[4]); not the headline workload")
    ap.add_argument("--tables", default="sharded", choices=["sharded", "replicated"], help="embedding-table layout when --gpus > 1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hbm", action="store_true", help="skip the HBM-bound kernel rooflines (shorter ncu launch lists)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--customers", type=int, default=V_CUSTOMERS, help="rows of the customer table (BASELINE configs[4]: 100000000)")
    ap.add_argument("--emb", type=int, default=JOINT, help="id-embedding and joint dimension (64 or 128)")
    ap.add_argument("--min-ms", type=float, default=1000.0, help="every timed region repeats its block of --steps steps until it covers this much time")
    ap.add_argument("--no-c3", action="store_true", help="skip the batch-65536 leg (BASELINE configs[2])")
    ap.add_argument("--no-big-index", action="store_true", help="skip the 10M / 100M-row index legs (BASELINE configs[3])")
    ap.add_argument("--big-index-rows", type=str, default="auto", help="comma list of corpus sizes for the row-sharded index leg (auto: 1e7, and 1e8 from 4 GPUs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-process multi-GPU parity checks")
    args = ap.parse_args()
    g = globals()
    g["V_CUSTOMERS"], g["E_ID"], g["JOINT"] = args.customers, args.emb, args.emb
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
