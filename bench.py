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
             OpenMP over all host cores) on a bounded sample of the same workload.  The reference's own binary
             cannot be built offline (un-vendored torchdrug), hence kind = "port".
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
def cpu_reference_sample(sample_batch, steps=1, warmup=0):
    """Time the C restatement of the reference's CPU rspmm (fwd + bwd) on `sample_batch` queries of the workload."""
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
    times = []
    for step in range(warmup + steps):
        start = time.perf_counter()
        output = cpu_ref.forward(csr, relation, input, "add", "mul")
        cpu_ref.backward(csr, relation, input, output, grad, "add", "mul")
        if step >= warmup:
            times.append(time.perf_counter() - start)
    seconds = sum(times) / len(times)
    gbs = edge_model_bytes(n, r, e, d, "both") / seconds / 1e9
    return {"value": gbs, "unit": UNIT, "cores": cpu_ref.num_threads(), "kind": "port",
            "sample": "same graph (N=%d, R'=%d, E=%d), %d of %d queries (D=%d), fwd+bwd sum x DistMult fp32, "
                      "%.2f s per pass; C restatement of torchdrug rspmm.cpp (OpenMP rows, atomic adds in backward) - "
                      "the reference's own binary is un-vendored and cannot be built offline"
                      % (n, r, e, sample_batch, BATCH, d, seconds),
            "seconds_per_step": seconds}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 8   # 8 of the 64 queries per step: ~1 s per step on 8 cores, so K + W steps end within minutes
    result = cpu_reference_sample(sample_batch, steps=max(args.steps, 1), warmup=args.warmup)
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
            "l2_policy": "inputs rotate over 2 operand sets of 484 MB each (> 126 MB L2) between timed steps"}


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
    e2e_ms, h2d, d2h = e2e_host_buffers(lib, edge_list, n, r, d, local_rank, max(min(args.steps, 5), 2), rank)
    del sets
    torch.cuda.empty_cache()
    queries = ultra_queries(device, rank, steps=max(min(args.steps, 5), 2))

    if world > 1:
        both = torch.tensor([elapsed_ms, e2e_ms, queries["ms_per_batch"]], device=device, dtype=torch.float64)
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, queries["ms_per_batch"] = both.tolist()

    if rank == 0:
        peak, peak_source = hbm_peak()
        ms_per_step = elapsed_ms / args.steps
        value = world * bytes_step / (ms_per_step * 1e-3) / 1e9
        achieved = bytes_fwd / (forward_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(),
            "pct_of_hbm_peak": 100.0 * value / world / peak,
            "edge_messages_per_s": world * 2.0 * e * d / (ms_per_step * 1e-3),
            "clocks": clocks,
            "e2e": {"value": world * bytes_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "numa_local_cores_per_rank": numa_cores,
                    "api": "ultra_rspmm_ctx_forward_backward (C ABI, pinned host buffers)"},
            "gpu_launches": launches,
            "ultra_queries": {"value": world * queries["queries_per_batch"] / (queries["ms_per_batch"] * 1e-3),
                              "unit": "queries/s", "ms_per_batch": queries["ms_per_batch"],
                              "batch_per_gpu": BATCH, "relation_graph_edges": queries["relation_graph_edges"],
                              "mode": queries["mode"],
                              "rspmm_edge_model_GBps": world * queries["rspmm_edge_model_bytes_per_batch"] / (queries["ms_per_batch"] * 1e-3) / 1e9,
                              "pct_of_hbm_peak": 100.0 * queries["rspmm_edge_model_bytes_per_batch"] / (queries["ms_per_batch"] * 1e-3) / 1e9 / peak,
                              "what": "ULTRA zero-shot tail+head ranking, 6+6 layers x 64-d, all entities as candidates, "
                                      "fresh batch per step, random-init weights, fp32 (TF32 off)"},
            "roofline": {"bound": "hbm", "kernel": "seg_reduce_kernel<float,4,add,mul> (forward)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_source, "algorithmic_bytes_per_launch": bytes_fwd,
                         "launch_ms": forward_ms,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         # companion figures SURVEY.md 8d asks for beside the edge-model fraction
                         "compulsory_bytes_per_launch": 4 * d * (2 * n + r) + 12 * e + 4 * (n + 1),
                         "compulsory_frac": (4 * d * (2 * n + r) + 12 * e + 4 * (n + 1)) / (forward_ms * 1e-3) / 1e9 / peak,
                         "ncu": {k: traffic[k] for k in ("l2_hit_rate_pct", "l1_hit_rate_pct", "l2_bytes_read_by_sm",
                                                         "l1tex_throughput_pct", "lts_throughput_pct", "source")
                                 if k in traffic} if traffic else None,
                         "binding_unit": "L1 data pipe (128 B/clk/SM: gathered row + relation row = 8 wavefronts per "
                                         "edge and 128-feature slab); L2->SM bandwidth for the relation-gradient pass",
                         "note": "edge-model bytes count one D-wide row gather per edge; slab-major scheduling "
                                 "serves those gathers from L2, so frac > 1 is expected and DRAM traffic is near the "
                                 "compulsory 2*N*D*4 bytes (see profiles/)"},
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
    ranker = nbf.UltraRanker(model.to(device).eval(), rel_model.to(device).eval(), graph)
    generator = torch.Generator().manual_seed(4096 + rank)
    batches = [triples[torch.randint(num_triple, (BATCH,), generator=generator)].to(device) for _ in range(warmup + steps)]
    try:
        predict, mode = ranker.capture(BATCH), "CUDA graph replay"     # launch overhead off the critical path
    except Exception:
        predict, mode = ranker.predict, "eager"
    with torch.no_grad():
        for batch in batches[:warmup]:
            predict(batch)
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for batch in batches[warmup:]:
            pred = predict(batch)     # a fresh batch every step (copied into the graph's input): nothing is cached
        stop.record()
        torch.cuda.synchronize()
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
