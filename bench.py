#!/usr/bin/env python
"""Benchmark of the rspmm hot path (BASELINE.json metric: "rspmm fwd+bwd GB/s (% HBM peak)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], C2): FB15k-237-shaped synthetic graph - 14,541 entities, 237 relations,
272,115 random triples + inverses (E = 544,230 before coalescing, R' = 474), query batch 64 folded into the
feature axis (D = 64 * 64 = 4,096 fp32 features), sum x DistMult, fp32.  One *step* = one forward + one backward
(gradients w.r.t. relation and input) of `generalized_rspmm` over that batch.

* `value`  : edge-model bytes (A = 4D(2E + 3N + 3R') + 24E + 8(N+1)) / device time per step, inputs
             resident in HBM, called through the public autograd operator.  Whole-job aggregate over N GPUs: every
             rank holds the full graph index and its own 64-query slab (weak scaling, no data-path collective).
* `e2e`    : the same metric through the torch-free C-ABI host-buffer call (`ultra_rspmm_ctx_forward_backward`):
             pinned HOST operands in, results copied back, H2D/D2H inside the timed region.
* `roofline`: forward kernel (`seg_reduce_kernel`, the dominant launch): algorithmic bytes A_f = 4D(E + R' + N) +
             12E + 4(N+1) per launch / its CUDA-event duration, against the measured HBM copy peak.
* `cpu_baseline` / `--impl reference`: the C restatement of torchdrug's compiled CPU rspmm (oracle/rspmm_cpu_ref.c,
             OpenMP over all host cores) on a bounded sample of the same workload (8 of the 64 queries per step).  The
             reference's own binary cannot be built offline (un-vendored torchdrug), hence kind = "port"; if
             `import torchdrug` works on the box, its real CPU operator is timed instead (kind = "reference binary").
* extra keys: `roofline_hbm_resident` (C4 and the 16.8 M-edge sweep shape, where the fraction of the HBM peak means what
             it says), `ceilings` (gather-bandwidth probes run on this box: L2 -> SM and HBM random rows), `c4_predict`
             (BASELINE configs[3]: YAGO3-10-shaped zero-shot evaluation, global batch 64 split over the ranks, NCCL
             gather of the ranks: strong scaling) and `c3_finetune` (configs[2]: CoDEx-L-shaped fine-tuning step, batch 64
             split over the ranks, NCCL gradient all-reduce).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rspmm fwd+bwd GB/s (% HBM peak)"
UNIT = "GB/s"
GRAPH = "fb15k237"
BATCH = 64
HIDDEN = 64
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback, used only when MEASURED_PEAKS.json is absent


def edge_model_bytes(n, r, e, d, which, elem=4):
    """Algorithmic bytes of SURVEY.md section 8d (int32 indices)."""
    idx = 12 * e + 4 * (n + 1)
    if which == "fwd":
        return elem * d * (e + r + n) + idx
    if which == "bwd":
        return elem * d * (e + 2 * n + 2 * r) + idx
    return edge_model_bytes(n, r, e, d, "fwd", elem) + edge_model_bytes(n, r, e, d, "bwd", elem)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as handle:
            return float(json.load(handle)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def build_hash():
    """Content hash of the CUDA sources the loaded library was built from (ties ncu captures to a build)."""
    try:
        from ultra_torchdrug_b200 import build
        return build.source_hash()[:16]
    except Exception:
        return None


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "dominant_kernel.json")) as handle:
            return json.load(handle)
    except Exception:
        return None


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.process = None
        self.thread = None
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.process = None

    def _read(self):
        for line in self.process.stdout:
            self.samples.append((time.time(), line.strip()))

    def count_between(self, start_time, stop_time):
        return sum(1 for t, _ in list(self.samples) if start_time <= t <= stop_time)

    def stop(self, start_time, stop_time):
        if self.process is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.process.terminate()
        try:
            self.process.wait(timeout=2)
        except Exception:
            self.process.kill()
        inside = [s for t, s in self.samples if start_time <= t <= stop_time + 0.15]
        clocks, max_clock, reasons = [], None, set()
        for sample in inside:
            fields = [f.strip() for f in sample.split(",")]
            if len(fields) < 6:
                continue
            try:
                clocks.append(float(fields[0]))
                max_clock = float(fields[1])
            except ValueError:
                continue
            for name, field in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), fields[2:6]):
                if field.lower().startswith("active"):
                    reasons.add(name)
        clocks.sort()
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": max_clock,
                "reasons": sorted(reasons), "samples": len(clocks)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------
REFERENCE_SAMPLE_BATCH = 8   # queries per step of the CPU arm: ~1 s per step on 8 cores, so K + W steps end within minutes


def torchdrug_operator():
    """The reference's real operator, if torchdrug is importable on this box (baseline/_ref or site-packages).
    It is not in this image (un-vendored, no wheel offline: BASELINE.md section 3), so this normally returns None and
    the CPU arm times the C restatement; the probe is here so that a box which has torchdrug times the real thing."""
    for extra in (os.path.join(ROOT, "baseline", "_ref"),):
        if os.path.isdir(extra) and extra not in sys.path:
            sys.path.append(extra)
    try:
        from torchdrug.layers import functional as td_functional
        if "ultra_torchdrug_b200" in (getattr(td_functional, "__file__", "") or ""):
            return None                      # our own import shim, not torchdrug
        return td_functional.generalized_rspmm
    except Exception:
        return None


def cpu_reference_sample(sample_batch, steps=1, warmup=0):
    """Time the reference's CPU rspmm (fwd + bwd) on `sample_batch` queries of the workload: torchdrug's own operator
    when importable, else the C restatement of it (oracle/rspmm_cpu_ref.c)."""
    import numpy as np
    from oracle import cpu_ref
    from ultra_torchdrug_b200 import synthetic

    cpu_ref.use_all_cores()
    edge_list, n, r = synthetic.named_graph(GRAPH)
    indices = edge_list[:, [1, 0, 2]].t().contiguous().numpy()
    values = np.ones(indices.shape[1], dtype=np.float32)
    csr = cpu_ref.CsrOperand(indices, values, (n, n, r))
    e = len(csr.col)
    d = sample_batch * HIDDEN
    rng = np.random.default_rng(1024)
    relation = rng.standard_normal((r, d), dtype=np.float32)
    input = rng.standard_normal((n, d), dtype=np.float32)
    grad = rng.standard_normal((n, d), dtype=np.float32)
    real = torchdrug_operator()
    if real is not None:
        import torch
        torch.set_num_threads(cpu_ref.num_threads())
        sparse = torch.sparse_coo_tensor(torch.from_numpy(indices), torch.from_numpy(values), (n, n, r))
        t_relation, t_input, t_grad = (torch.from_numpy(x) for x in (relation, input, grad))
    times = []
    for step in range(warmup + steps):
        start = time.perf_counter()
        if real is not None:
            a, b = t_relation.clone().requires_grad_(), t_input.clone().requires_grad_()
            real(sparse, a, b, sum="add", mul="mul").backward(t_grad)
        else:
            output = cpu_ref.forward(csr, relation, input, "add", "mul")
            cpu_ref.backward(csr, relation, input, output, grad, "add", "mul")
        if step >= warmup:
            times.append(time.perf_counter() - start)
    seconds = sum(times) / len(times)
    gbs = edge_model_bytes(n, r, e, d, "both") / seconds / 1e9
    what = ("torchdrug.layers.functional.generalized_rspmm on CPU (the reference's own binary)" if real is not None else
            "C restatement of torchdrug rspmm.cpp (OpenMP rows, atomic adds in backward) - the reference's own binary is "
            "un-vendored and cannot be built offline; `import torchdrug` was probed and is absent")
    return {"value": gbs, "unit": UNIT, "cores": cpu_ref.num_threads(), "kind": "reference binary" if real is not None else "port",
            "sample": "same graph (N=%d, R'=%d, E=%d), %d of %d queries (D=%d of %d), fwd+bwd sum x DistMult fp32, "
                      "%.2f s per pass; %s" % (n, r, e, sample_batch, BATCH, d, BATCH * HIDDEN, seconds, what),
            "seconds_per_step": seconds}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    result = cpu_reference_sample(REFERENCE_SAMPLE_BATCH, steps=max(args.steps, 1), warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": result["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": result["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {k: result[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": result["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config():
    return {"workload": "C2 FB15k-237-shaped synthetic graph: 14541 entities, 237 relations (+inverses = 474), "
                        "272115 triples + inverses, query batch 64 x 64-d folded into D=4096 fp32 features, "
                        "sum x DistMult, one step = rspmm forward + backward (grad relation + grad input)",
            "graph": GRAPH, "batch_per_gpu": BATCH, "dim": BATCH * HIDDEN,
            "l2_policy": "inputs rotate over 2 operand sets of 484 MB each (> 126 MB L2) between timed steps",
            "reference_arm_sample": "--impl reference (CPU) runs %d of the %d queries per step (D = %d of %d): a bounded "
                                    "sample of this workload; GB/s normalises the width" % (
                                        REFERENCE_SAMPLE_BATCH, BATCH, REFERENCE_SAMPLE_BATCH * HIDDEN, BATCH * HIDDEN)}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index):
    """Multi-GPU runs: keep this rank (and the pinned host buffers it allocates for the e2e leg) on the CPU cores
    that are NUMA-local to its GPU, so that 8 ranks do not push their PCIe traffic across the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def gather_ceilings(lib, device, rows, dim):
    """What a gather of 512-byte row pieces can reach on THIS box (`ultra_probe_gather`: the rspmm access pattern with
    no ids and no arithmetic): out of L2 for the column slab of a (rows, dim) matrix - the C2 regime - and out of HBM
    for random rows of a 4 GiB buffer - the regime of graphs whose slab exceeds L2."""
    import torch
    results = {}
    sink = torch.zeros(1, device=device)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, n_rows, stride in (("l2_to_sm", rows, 4 * dim), ("hbm_random_rows", (4 << 30) // 512, 512)):
        buffer = torch.zeros(n_rows * stride // 4, device=device)
        moved = ctypes.c_int64()
        blocks, iters = 148 * 8, 1024

        def launch():
            status = lib.ultra_probe_gather(buffer.data_ptr(), n_rows, stride, iters, blocks, sink.data_ptr(),
                                            ctypes.byref(moved), stream)
            if status:
                raise RuntimeError("ultra_probe_gather failed with %d" % status)

        for _ in range(2):
            launch()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(5):
            launch()
        stop.record()
        torch.cuda.synchronize()
        results[name] = {"GBps": 5 * moved.value / (start.elapsed_time(stop) * 1e-3) / 1e9,
                         "footprint_MB": n_rows * 512 / 1e6, "row_stride_bytes": stride}
        del buffer
    torch.cuda.empty_cache()
    results["how"] = ("ultra_probe_gather: 148 x 8 CTAs x 8 warps, 4 LDG.128 in flight per lane, pseudo-random 512-byte "
                      "row pieces, L1::no_allocate; measured in this run")
    return results


def committed_traffic(name):
    """DRAM bytes per launch from the committed ncu capture of this shape (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as handle:
            return json.load(handle).get(name)
    except Exception:
        return None


def hbm_resident_rooflines(device, peak):
    """Shapes whose gathered slab does not fit the 126 MB L2, where the fraction of the measured HBM peak means what it
    says: C4 (YAGO3-10 shape, 63 MB slab + result rows + ids) and the largest shape of the configs[4] sweep
    (16.8 M edges, N = E / 32, R' = 474, D = 4096: 268 MB slab)."""
    import torch
    from ultra_torchdrug_b200 import functional as F, synthetic

    def measure(index, n, r, d, iters):
        generator = torch.Generator(device=device).manual_seed(5)
        relation = torch.randn(r, d, device=device, generator=generator)
        input = torch.randn(n, d, device=device, generator=generator)
        grad = torch.randn(n, d, device=device, generator=generator)
        out = index.forward(relation, input)
        index.backward(relation, input, None, grad)
        torch.cuda.synchronize()
        events = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        events[0].record()
        for _ in range(iters):
            out = index.forward(relation, input)
        events[1].record()
        for _ in range(iters):
            index.backward(relation, input, None, grad)
        events[2].record()
        torch.cuda.synchronize()
        return events[0].elapsed_time(events[1]) / iters, events[1].elapsed_time(events[2]) / iters

    entries = []
    shapes = [("C4 YAGO3-10 shape", "yago310", None), ("C5 sweep 16.8M edges", None, (1 << 24, 1 << 19, 474))]
    for label, graph, uniform in shapes:
        if graph:
            edge_list, n, r = synthetic.named_graph(graph)
            indices = edge_list[:, [1, 0, 2]].t().contiguous().to(device)
        else:
            e_raw, n, r = uniform
            generator = torch.Generator(device=device).manual_seed(1024)
            indices = torch.stack([torch.randint(n, (e_raw,), device=device, generator=generator),
                                   torch.randint(n, (e_raw,), device=device, generator=generator),
                                   torch.randint(r, (e_raw,), device=device, generator=generator)])
        index = F.GraphIndex(indices, torch.ones(indices.shape[1], device=device), (n, n, r))
        d = BATCH * HIDDEN
        forward_ms, backward_ms = measure(index, n, r, d, 3)
        e = index.nnz
        for which, ms, passes in (("forward", forward_ms, "fwd"), ("forward + backward", forward_ms + backward_ms, "both")):
            achieved = edge_model_bytes(n, r, e, d, passes) / (ms * 1e-3) / 1e9
            entries.append({"shape": "%s: N=%d R'=%d E=%d D=%d, gathered slab %d MB" % (label, n, r, e, d, n * 512 // 1000000),
                            "pass": which, "bound": "hbm", "ms": ms, "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "algorithmic_bytes": edge_model_bytes(n, r, e, d, passes),
                            "gathers_per_edge": 1 if passes == "fwd" else 4,
                            "traffic": committed_traffic("%s %s" % (label.split()[0], passes))})
        del index, indices
        torch.cuda.empty_cache()
    return entries


def host_copy_ceiling(device, h2d_bytes, d2h_bytes, steps, barrier):
    """What the host side of this box allows for the e2e leg: the same byte volumes as plain large pinned copies, H2D and
    D2H concurrently on two streams (PCIe is full duplex), every rank at the same time."""
    import torch
    up_host = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    down_host = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    up_dev = torch.empty(h2d_bytes, dtype=torch.uint8, device=device)
    down_dev = torch.empty(d2h_bytes, dtype=torch.uint8, device=device)
    streams = [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)]

    def once():
        with torch.cuda.stream(streams[0]):
            up_dev.copy_(up_host, non_blocking=True)
        with torch.cuda.stream(streams[1]):
            down_host.copy_(down_dev, non_blocking=True)
        streams[0].synchronize()
        streams[1].synchronize()

    once()
    barrier()
    start = time.perf_counter()
    for _ in range(steps):
        once()
    seconds = (time.perf_counter() - start) / steps
    return seconds * 1e3


def c4_predict(device, rank, world, steps, warmup=2):
    """BASELINE.json configs[3]: ULTRA zero-shot evaluation on the YAGO3-10-shaped graph with the GLOBAL batch of 64
    test triples split over the ranks (strong scaling: 64 / 32 / 16 / 8 triples per GPU at 1 / 2 / 4 / 8 GPUs).  One step =
    `predict` of this rank's triples (relation model + tail pass + head pass, all 123,182 entities as candidates) +
    filtered ranks on the device + one NCCL all-gather of the (B, 2) ranks (reference ultra/engine.py:130-150)."""
    import torch
    from ultra_torchdrug_b200 import nbf, sharding, synthetic, task
    from ultra_torchdrug_b200.compat.torchdrug import data

    num_node, num_relation, num_triple = synthetic.SHAPES["yago310"]
    triples = synthetic.triples(num_node, num_relation, num_triple)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
    torch.manual_seed(1024)
    model, rel_model = nbf.ultra_models(num_relation)
    ranker = nbf.UltraRanker(model.to(device).eval(), rel_model.to(device).eval(), graph)
    evaluator = task.ShardedEvaluator(ranker, rank=rank, world_size=world)
    start_q, stop_q = sharding.query_slab(BATCH, rank, world)
    generator = torch.Generator().manual_seed(8192)            # the same global batches on every rank
    batches = [triples[torch.randint(num_triple, (BATCH,), generator=generator)].to(device) for _ in range(warmup + steps)]
    try:
        predict, mode = ranker.capture(stop_q - start_q), "CUDA graph replay"
    except Exception:
        predict, mode = ranker.predict, "eager"
    with torch.no_grad():
        for batch in batches[:warmup]:
            ranks = evaluator(batch, predict)
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for batch in batches[warmup:]:
            ranks = evaluator(batch, predict)
        stop.record()
        torch.cuda.synchronize()
    mrr = float((1.0 / ranks.float()).mean())
    with torch.no_grad():
        spread = float(predict(batches[-1][start_q:stop_q]).float().std())   # random weights: says whether scores vary at all
    result = {"ms_per_global_batch": start.elapsed_time(stop) / steps, "global_batch": BATCH,
              "batch_per_gpu": stop_q - start_q, "mode": mode, "relation_graph_edges": int(ranker.rel_graph.num_edge),
              "mrr_last_batch_random_weights": mrr, "score_std_last_batch": spread}
    del ranker, evaluator, predict, graph
    torch.cuda.empty_cache()
    return result


def c3_finetune(device, rank, world, steps, warmup=2):
    """BASELINE.json configs[2]: one ULTRA fine-tuning step on the CoDEx-L-shaped graph, global batch 64 triples x (1 + 128
    strict negatives) split over the ranks, forward + backward + NCCL all-reduce of the 0.78 MB of gradients + AdamW
    (reference ultra/engine.py:48-86, ultra/task.py:160-195, 264-275)."""
    import torch
    from ultra_torchdrug_b200 import functional as F, nbf, synthetic, task
    from ultra_torchdrug_b200.compat.torchdrug import data

    num_node, num_relation, num_triple = synthetic.SHAPES["codex_l"]
    triples = synthetic.triples(num_node, num_relation, num_triple)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
    torch.manual_seed(1024)                                     # same initial weights on every rank
    model, rel_model = nbf.ultra_models(num_relation)
    step = task.FinetuneStep(model.to(device).train(), rel_model.to(device).train(), graph, rank=rank, world_size=world)
    generator = torch.Generator().manual_seed(7)                # the same global batches on every rank
    batches = [triples[torch.randint(num_triple, (BATCH,), generator=generator)].to(device) for _ in range(warmup + steps)]
    torch.manual_seed(4096 + rank)                              # negatives differ per rank
    # No eager backward may run on the default stream before the capture: autograd would bind the parameters' AccumulateGrad
    # nodes to that stream and the capture (which warms up on a side stream) would be invalidated.
    mode, launches = "eager", F.launch_count()
    if os.environ.get("ULTRA_BENCH_FINETUNE_GRAPH", "1") != "0":
        step.capture(BATCH)                                     # the whole step (incl. the NCCL all-reduce) as one CUDA graph
        mode = "CUDA graph replay"
        launches_per_step = (F.launch_count() - launches) / 4   # 3 warm-up steps + the captured one
    else:
        step(batches[0])
        launches_per_step = F.launch_count() - launches
    for batch in batches[:warmup]:
        step(batch)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for batch in batches[warmup:]:
        loss = step(batch)
    stop.record()
    torch.cuda.synchronize()
    result = {"ms_per_step": start.elapsed_time(stop) / steps, "global_batch": BATCH, "negatives": step.num_negative,
              "library_launches_per_step": launches_per_step, "loss_share_last_step": float(loss), "mode": mode,
              "gradient_bytes": 4 * sum(p.numel() for p in step.parameters)}
    del step, graph
    torch.cuda.empty_cache()
    return result


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rspmm path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    numa_cores = None
    if world > 1:
        numa_cores = bind_to_gpu_numa_node(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False   # reference run_full.py:19-20
    torch.backends.cudnn.allow_tf32 = False

    from ultra_torchdrug_b200 import _lib, functional as F, synthetic

    lib = _lib.lib()
    edge_list, n, r = synthetic.named_graph(GRAPH)
    sparse = synthetic.operator_operand(edge_list, n, r, device)
    index = F.graph_index(sparse)
    e = index.nnz
    d = BATCH * HIDDEN
    generator = torch.Generator(device=device).manual_seed(1024 + rank)
    sets = []
    for _ in range(2):
        sets.append({
            "relation": torch.randn(r, d, device=device, generator=generator).requires_grad_(),
            "input": torch.randn(n, d, device=device, generator=generator).requires_grad_(),
            "grad": torch.randn(n, d, device=device, generator=generator),
        })
    bytes_fwd = edge_model_bytes(n, r, e, d, "fwd")
    bytes_step = edge_model_bytes(n, r, e, d, "both")

    forward_events = []

    def step(i, record=False):
        operands = sets[i % 2]
        operands["relation"].grad = None
        operands["input"].grad = None
        if record:
            begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            begin.record()
        output = F.generalized_rspmm(sparse, operands["relation"], operands["input"], sum="add", mul="mul")
        if record:
            end.record()
            forward_events.append((begin, end))
        output.backward(operands["grad"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None      # nvidia-smi needs ~0.1 s to start: before the warm-up
    for i in range(args.warmup):
        step(i)
    barrier()
    launches_before = lib.ultra_rspmm_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall_start = time.time()
    start.record()
    for i in range(args.steps):
        step(i, record=True)
    stop.record()
    barrier()
    wall_stop = time.time()
    launches = int(lib.ultra_rspmm_launch_count() - launches_before)
    elapsed_ms = start.elapsed_time(stop)
    clocks = None
    if sampler:
        window, where = (wall_start, wall_stop), "timed region"
        if sampler.count_between(*window) < 3:
            # the timed region is shorter than a few 20 ms sampling periods: keep the very same steps running (untimed)
            # until nvidia-smi has reported three samples under that load
            deadline = time.time() + 5.0
            while sampler.count_between(wall_stop, time.time()) < 3 and time.time() < deadline:
                for i in range(args.steps):
                    step(i)
                torch.cuda.synchronize()
            window, where = (wall_start, time.time()), "timed region + the same steps repeated right after it"
        clocks = sampler.stop(*window)
        clocks["window"] = where
    forward_ms = sum(b.elapsed_time(e_) for b, e_ in forward_events) / max(len(forward_events), 1)

    # ---- end to end through the C ABI with host buffers (H2D + D2H inside the timed region) -------
    del sets
    torch.cuda.empty_cache()
    barrier()                                                   # every rank loads the host side at the same time
    e2e_steps = max(min(args.steps, 5), 2)
    e2e_ms, h2d, d2h = e2e_host_buffers(lib, edge_list, n, r, d, local_rank, e2e_steps, rank)
    barrier()
    copy_ms = host_copy_ceiling(device, h2d, d2h, e2e_steps, barrier)
    queries = ultra_queries(device, rank, steps=max(min(args.steps, 5), 2))
    extra_steps = max(min(args.steps, 5), 3)
    barrier()
    c4 = c4_predict(device, rank, world, extra_steps)
    barrier()
    try:
        c3 = c3_finetune(device, rank, world, extra_steps)
    except Exception as error:                                  # noqa: BLE001 - never lose the headline line to an extra leg
        if world > 1:
            raise                                               # ranks must not diverge around a collective
        sys.stderr.write("c3_finetune leg failed: %r\n" % (error,))
        c3 = {"ms_per_step": float("nan"), "gradient_bytes": 0, "error": repr(error)}

    if world > 1:
        both = torch.tensor([elapsed_ms, e2e_ms, queries["ms_per_batch"], copy_ms, c4["ms_per_global_batch"], c3["ms_per_step"]],
                            device=device, dtype=torch.float64)
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, queries["ms_per_batch"], copy_ms, c4["ms_per_global_batch"], c3["ms_per_step"] = both.tolist()

    if rank == 0:
        peak, peak_source = hbm_peak()
        ceilings = gather_ceilings(lib, device, n, d)
        resident = hbm_resident_rooflines(device, peak) if world == 1 else None
        ms_per_step = elapsed_ms / args.steps
        value = world * bytes_step / (ms_per_step * 1e-3) / 1e9
        achieved = bytes_fwd / (forward_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        l1_peak = 148 * 128 * sm_mhz * 1e6 / 1e9                  # GB/s: 128 B per clock and SM through the L1 data pipe
        l1_bytes = 2 * 4 * d * e                                  # gathered row + relation row of every edge cross it
        l2_peak = ceilings["l2_to_sm"]["GBps"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(),
            "pct_of_hbm_peak": 100.0 * value / world / peak,
            "edge_messages_per_s": world * 2.0 * e * d / (ms_per_step * 1e-3),
            "clocks": clocks,
            "e2e": {"value": world * bytes_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "numa_local_cores_per_rank": numa_cores,
                    "api": "ultra_rspmm_ctx_forward_backward (C ABI, pinned host buffers)",
                    "host_copy_ceiling": {"ms_per_step": copy_ms, "GBps_each_way": h2d / (copy_ms * 1e-3) / 1e9,
                                          "what": "the same %d + %d bytes as plain pinned copies, H2D and D2H concurrently, all "
                                                  "%d ranks at once (max over ranks)" % (h2d, d2h, world)},
                    "frac_of_host_copy_ceiling": copy_ms / e2e_ms},
            "gpu_launches": launches,
            "ultra_queries": {"value": world * queries["queries_per_batch"] / (queries["ms_per_batch"] * 1e-3),
                              "unit": "queries/s", "ms_per_batch": queries["ms_per_batch"],
                              "batch_per_gpu": BATCH, "relation_graph_edges": queries["relation_graph_edges"],
                              "relation_graph_note": "uniform synthetic triples make the graph of relations fully dense "
                                                     "(4 * 474^2 edges); its 6 passes are an artifact-sized share of the batch",
                              "mode": queries["mode"], "scaling": "weak",
                              "rspmm_edge_model_GBps": world * queries["rspmm_edge_model_bytes_per_batch"] / (queries["ms_per_batch"] * 1e-3) / 1e9,
                              "pct_of_hbm_peak": 100.0 * queries["rspmm_edge_model_bytes_per_batch"] / (queries["ms_per_batch"] * 1e-3) / 1e9 / peak,
                              "what": "ULTRA zero-shot tail+head ranking: scores of all entities + filtered ranks on the device "
                                      "(the reference copies the (B, 2, N) scores to the host instead, task.py:261-263), 6+6 layers x "
                                      "64-d, fresh batch per step, random-init weights, fp32 (TF32 off)"},
            "c4_predict": dict(c4, metric="ULTRA queries/s (BASELINE configs[3], YAGO3-10 shape)", unit="queries/s",
                               value=2 * BATCH / (c4["ms_per_global_batch"] * 1e-3), scaling="strong", n_gpus=world,
                               collective="NCCL all_gather of the (B, 2) filtered ranks" if world > 1 else "none (1 rank)",
                               what="predict + filtered ranks on device + gather, global batch 64 triples = 128 queries"),
            "c3_finetune": dict(c3, metric="ULTRA fine-tuning triples/s (BASELINE configs[2], CoDEx-L shape)", unit="triples/s",
                                value=BATCH / (c3["ms_per_step"] * 1e-3), scaling="strong", n_gpus=world,
                                collective="NCCL all_reduce of %d gradient bytes" % c3["gradient_bytes"] if world > 1 else "none (1 rank)",
                                what="strict negatives + fwd + bwd + all-reduce + AdamW, global batch 64 x (1 + 128)"),
            "ceilings": ceilings,
            "roofline": {"kernel": "seg_reduce_kernel<float,4,add,mul> (forward, C2: the dominant launch of the step)",
                         "bound": "l2->sm",
                         "bound_note": "NOT hbm: the 7.4 MB column slab of the gathered operand is L2-resident (slab-major "
                                       "schedule), DRAM moves about the compulsory bytes (`traffic`); the gathers are bounded by "
                                       "L2 -> SM bandwidth and by the L1 data pipe both operands cross",
                         "achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                         "peak_source": "L2 -> SM gather ceiling measured in this run (`ceilings.l2_to_sm`)",
                         "algorithmic_bytes_per_launch": bytes_fwd, "launch_ms": forward_ms,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": (traffic or {}).get("source"),
                         "l1_data_pipe": {"bytes_per_launch": l1_bytes, "achieved": l1_bytes / (forward_ms * 1e-3) / 1e9,
                                          "peak": l1_peak, "frac": l1_bytes / (forward_ms * 1e-3) / 1e9 / l1_peak,
                                          "peak_source": "148 SMs x 128 B/clk x %.0f MHz (clock sampled in this run)" % sm_mhz},
                         "hbm_edge_model": {"achieved": achieved, "peak": peak, "frac": achieved / peak, "peak_source": peak_source,
                                            "note": "edge-model bytes over the measured HBM copy peak; > 1 because the gathers "
                                                    "are L2 hits - it is NOT an HBM roofline (see roofline_hbm_resident for "
                                                    "shapes where it is)"},
                         # companion figures SURVEY.md 8d asks for beside the edge-model fraction
                         "compulsory_bytes_per_launch": 4 * d * (2 * n + r) + 12 * e + 4 * (n + 1),
                         "compulsory_frac_of_hbm": (4 * d * (2 * n + r) + 12 * e + 4 * (n + 1)) / (forward_ms * 1e-3) / 1e9 / peak,
                         "ncu": {k: traffic[k] for k in ("l2_hit_rate_pct", "l1_hit_rate_pct", "l2_bytes_read_by_sm",
                                                         "l1tex_throughput_pct", "lts_throughput_pct",
                                                         "relation_table_l2_hit_rate_pct", "library_source_hash")
                                 if k in traffic} if traffic else None},
            "roofline_hbm_resident": resident,
            "library_source_hash": build_hash(),
        }
        if world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_sample(16, steps=4, warmup=1).items()
                                    if k != "seconds_per_step"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ultra_queries(device, rank, steps, warmup=2):
    """ULTRA zero-shot ranking throughput (BASELINE.json second metric): one step = `predict` on a fresh batch of
    64 triples = 128 ranking queries (tail + head) = 1 relation-graph pass + 2 entity-graph passes of 6 layers each
    (reference ultra/task.py:228-263), random-init weights of the shipped architecture, fp32, TF32 off."""
    import torch
    from ultra_torchdrug_b200 import nbf, synthetic
    from ultra_torchdrug_b200.compat.torchdrug import data

    num_node, num_relation, num_triple = synthetic.SHAPES[GRAPH]
    triples = synthetic.triples(num_node, num_relation, num_triple)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
    torch.manual_seed(1024)
    model, rel_model = nbf.ultra_models(num_relation)
    from ultra_torchdrug_b200 import task
    ranker = nbf.UltraRanker(model.to(device).eval(), rel_model.to(device).eval(), graph)
    evaluator = task.ShardedEvaluator(ranker)                           # this rank's own batch: no gather (weak scaling)
    generator = torch.Generator().manual_seed(4096 + rank)
    batches = [triples[torch.randint(num_triple, (BATCH,), generator=generator)].to(device) for _ in range(warmup + steps)]
    try:
        predict, mode = ranker.capture(BATCH), "CUDA graph replay"     # launch overhead off the critical path
    except Exception:
        predict, mode = ranker.predict, "eager"
    with torch.no_grad():
        for batch in batches[:warmup]:
            evaluator.local_ranks(batch, predict)
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for batch in batches[warmup:]:
            # a fresh batch every step (copied into the graph's input): nothing is cached; scores of all entities for the
            # 128 queries, then the filtered ranks of the true tails / heads on the device (reference task.py:279-315)
            ranks = evaluator.local_ranks(batch, predict)
        stop.record()
        torch.cuda.synchronize()
        pred = predict(batches[-1])
    ms = start.elapsed_time(stop) / steps
    # edge-model bytes of the 18 operator calls of one batch: 6 relation-graph layers + 2 x 6 entity-graph layers
    from ultra_torchdrug_b200 import functional as F
    d = BATCH * HIDDEN
    entity_index = F.graph_index(graph.undirected(add_inverse=True).adjacency.transpose(0, 1))
    relation_index = F.graph_index(ranker.rel_graph.adjacency.transpose(0, 1))
    model_bytes = 12 * edge_model_bytes(num_node, 2 * num_relation, entity_index.nnz, d, "fwd") + \
        6 * edge_model_bytes(2 * num_relation, 4, relation_index.nnz, d, "fwd")
    return {"ms_per_batch": ms, "queries_per_batch": 2 * BATCH, "relation_graph_edges": int(ranker.rel_graph.num_edge),
            "mode": mode,
            "rspmm_edge_model_bytes_per_batch": model_bytes, "score_checksum": float(pred.float().mean())}


def e2e_host_buffers(lib, edge_list, n, r, d, device_index, steps, rank):
    """fwd + bwd through `ultra_rspmm_ctx_forward_backward` with pinned host operands; returns (ms/step, h2d, d2h)."""
    import numpy as np
    import torch
    from ultra_torchdrug_b200 import _lib

    indices = edge_list[:, [1, 0, 2]].t().contiguous().numpy()
    values = np.ones(indices.shape[1], dtype=np.float32)
    ctx = ctypes.c_void_p()
    _lib.check(lib.ultra_rspmm_ctx_create(ctypes.byref(ctx), device_index), "ultra_rspmm_ctx_create")
    try:
        _lib.check(lib.ultra_rspmm_ctx_set_graph(ctx, indices.ctypes.data, values.ctypes.data, indices.shape[1], n, n, r,
                                                 _lib.F32), "ultra_rspmm_ctx_set_graph")
        generator = torch.Generator().manual_seed(2048 + rank)
        relation = torch.randn(r, d, generator=generator).pin_memory()
        input = torch.randn(n, d, generator=generator).pin_memory()
        grad = torch.randn(n, d, generator=generator).pin_memory()
        output = torch.empty(n, d).pin_memory()
        grad_relation = torch.empty(r, d).pin_memory()
        grad_input = torch.empty(n, d).pin_memory()

        def call():
            _lib.check(lib.ultra_rspmm_ctx_forward_backward(
                ctx, relation.data_ptr(), input.data_ptr(), grad.data_ptr(), output.data_ptr(),
                grad_relation.data_ptr(), grad_input.data_ptr(), d, 0, 0), "ultra_rspmm_ctx_forward_backward")

        call()
        call()
        start = time.perf_counter()
        for _ in range(steps):
            call()   # synchronises before returning
        seconds = (time.perf_counter() - start) / steps
        h2d = (relation.numel() + input.numel() + grad.numel()) * 4
        d2h = (output.numel() + grad_relation.numel() + grad_input.numel()) * 4
        return seconds * 1e3, h2d, d2h
    finally:
        lib.ultra_rspmm_ctx_destroy(ctx)


def main():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
