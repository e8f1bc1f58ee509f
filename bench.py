#!/usr/bin/env python
"""bench.py — HeteroGIN train-step throughput on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfgC|cfgA|cfgD|qt|gat] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic datanet-shaped samples:
CSR build -> L heterogeneous GIN layers -> readout -> sqrt(MAPE) -> backward -> Adam
(train.py:27-44).  Prints ONE JSON line on rank 0.

  value     graphs/s with the batch already resident in HBM when the timed region starts
  e2e       graphs/s through the public API with HOST (pinned) batches: H2D of the step's
            inputs + step + D2H read of the loss inside the timed region
  roofline  live CUDA-event timing of the aggregation kernel (hgin_gin_combine) over the timed
            region against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's CPU path on this box's host cores

`--impl reference` times the reference's own CPU implementation of the path (the oracle port:
the reference is pure Python on PyG, which cannot be installed here — DESIGN.md) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # config.json defaults (emb 8, 1 layer, readout [128,32]) at the reference's train batch of 8
    "cfgA": dict(batch=8, emb=8, layers=1, mlp=[128, 32], desc="configs[1]: config.json model, batch 8 of 50-node topologies"),
    # the large single-GPU configuration: 1024 topologies / step, hidden 128, 4 GIN layers
    "cfgC": dict(batch=1024, emb=128, layers=4, mlp=[128, 32], desc="configs[2]: 1024 topologies/step, hidden 128, 4 GIN layers"),
}
METRIC = "HeteroGIN train graphs/sec"


def model_kwargs(w):
    return dict(node_embedding_size=w["emb"], message_passing_layers=w["layers"], dropout=0.0, concat_path=True,
                bl_features=False, divided_features=False, global_feats=False, mlp_layers=list(w["mlp"]),
                act="torch.nn.PReLU()", mlp_head_act=None, mlp_bn=False)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """`nvidia-smi -lms 100` running beside the timed region: SM clock and throttle reasons."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        time.sleep(0.25)   # let the first sample land before the timed region starts

    def result(self):
        samples = []
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out, _ = self.proc.communicate()
            samples = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.count(",") >= 5]
        sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(samples)}


ALL_SAMPLES = []   # the samples behind the host batches (the device-resident dataset arm reuses them)


def make_host_batches(w, count, rank, pin, collate="csr"):
    from gnn_link_prediction_b200.data import Batch, CONV_EDGE_TYPES
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    ds = SyntheticDataset(w["batch"] * count, num_topologies=16, seed=1997 + 100003 * rank)
    batches = []
    for b in range(count):
        samples = [ds[b * w["batch"] + i] for i in range(w["batch"])]
        ALL_SAMPLES.extend(samples)
        # default: per-sample CSRs (built once per sample by K0, cached) are concatenated by the collate, so
        # the step runs no CSR build; --collate coo ships the COO lists and K0 runs inside every step
        batch = Batch.from_data_list(samples, index_dtype=torch.int32, edge_types=CONV_EDGE_TYPES, batch_vector=False,
                                     csr=collate == "csr", keep_coo=collate != "csr")
        batches.append(batch.pin_memory() if pin else batch)
    return batches


def batch_counts(batch):
    from gnn_link_prediction_b200.data import CONV_EDGE_TYPES
    edges = sum((batch[et]["csr_dst_col"].shape[0] if "csr_dst_col" in batch[et] else batch[et].edge_index.shape[1])
                for et in CONV_EDGE_TYPES)
    return batch.num_graphs, edges


def copy_batch_to_device(host):
    """What a user does per step: `sample.cuda()` (train.py:28) — a fresh device copy of every tensor."""
    from gnn_link_prediction_b200.data import Batch
    dev = Batch()
    for nt in host.node_types:
        for k, v in host[nt].items():
            dev[nt][k] = v.cuda(non_blocking=True)
    for et in host.edge_types:
        for k, v in host[et].items():
            dev[et][k] = v.cuda(non_blocking=True)
    dev.__dict__["num_graphs"] = host.num_graphs
    return dev


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


CPU_SAMPLE = {"cfgA": (8, 40), "cfgC": (256, 4)}     # (topologies per step, timed steps) of the CPU arms


def cpu_reference_run(w, sample_graphs, steps, warmup, threads=None):
    """The reference's CPU path (oracle port) on a bounded sample: `sample_graphs` topologies per step."""
    from oracle import hgin_oracle
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    # all host cores this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently time the CPU arm on ONE thread
    torch.set_num_threads(threads or host_cores())
    cores = torch.get_num_threads()
    ds = SyntheticDataset(sample_graphs, num_topologies=min(16, sample_graphs))
    batch = Batch.from_data_list([ds[i] for i in range(sample_graphs)])
    torch.manual_seed(1997)
    model = hgin_oracle.HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **model_kwargs(w))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0)
    model.train()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        hgin_oracle.train_step(model, opt, batch)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": sample_graphs * steps / total, "unit": "graphs/s", "cores": cores, "kind": "port",
            "sample": f"{sample_graphs} topologies/step x {steps} steps (+{warmup} warm-up), fwd+bwd+Adam, "
                      f"torch {torch.__version__} CPU, {cores} threads of {os.cpu_count()} cores",
            "ms_per_step": 1e3 * total / steps}


def run_reference(args, w, rank):
    if rank != 0:
        return
    # a bounded sample of the workload: cfgA is the reference's own batch of 8; cfgC runs 256 of the 1024
    # topologies per step (~2 s per step on 16 host threads; graphs/s is flat in the batch size from 128 up,
    # DESIGN.md), so a few steps give a stable number.  ALL host cores, also under torchrun.
    sample, max_steps = CPU_SAMPLE[args.workload]
    steps = max(1, min(args.steps, max_steps))
    r = cpu_reference_run(w, sample, steps, min(args.warmup, 1), threads=host_cores())
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "graphs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": w["desc"], "emb": w["emb"], "layers": w["layers"],
                       "batch_per_step": sample},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def timed_steps(step_fn, steps, flush_buf):
    """K steps, each bracketed by its own CUDA events on the current stream; with `flush_buf` the
    L2 is flushed (a >L2 buffer is overwritten) between steps, outside the per-step interval."""
    evs = []
    for i in range(steps):
        if flush_buf is not None:
            flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_fn(i)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfgC", choices=sorted(WORKLOADS) + ["cfgD", "qt", "gat"])
    ap.add_argument("--math", default=None, choices=["fp32", "tf32", "bf16"],
                    help="dense-layer arithmetic; default tf32 tensor cores for cfgC (BASELINE configs[2] allows "
                         "reduced-precision MLP GEMMs), fp32 for cfgA")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfgA line reported beside the default cfgC run")
    ap.add_argument("--collate", default="csr", choices=["csr", "coo"],
                    help="csr: the collate concatenates cached per-sample CSRs (no CSR build in the step); "
                         "coo: ship COO edge lists, hgin_csr_build runs inside every step")
    ap.add_argument("--stage", default="packed", choices=["packed", "tensors"],
                    help="e2e arm: H2D of one packed pinned buffer per batch (data.pack_batch) or one copy per tensor")
    ap.add_argument("--readback", default="deferred", choices=["deferred", "sync"],
                    help="e2e arm: how each step's loss reaches the host (see e2e_run)")
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="replay the step as a CUDA graph (default: on for cfgA at 1 GPU)")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload == "cfgD":
        import bench_conv
        return bench_conv.main(args)
    if args.workload == "qt":
        import bench_qt
        return bench_qt.main(args)
    if args.workload == "gat":
        import bench_gat
        return bench_gat.main(args)
    w = WORKLOADS[args.workload]
    if args.math is None:
        args.math = "tf32" if args.workload == "cfgC" else "fp32"
    if args.impl == "reference":
        return run_reference(args, w, rank)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    from gnn_link_prediction_b200.parallel import Communicator
    comm = Communicator.from_env("nccl")
    from gnn_link_prediction_b200 import ops
    from gnn_link_prediction_b200.models import HetroGIN, MATH_BF16, MATH_FP32, MATH_TF32
    from gnn_link_prediction_b200.profiling import KernelTimer
    from gnn_link_prediction_b200.train import TrainStep

    args.warmup = max(args.warmup, 3)
    n_host = 2
    host = make_host_batches(w, n_host, rank, pin=True, collate=args.collate)
    graphs, edges = batch_counts(host[0])
    h2d_bytes = host[0].nbytes()

    torch.manual_seed(1997)
    model = HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **model_kwargs(w)).cuda().train()
    model.set_math_mode({"tf32": MATH_TF32, "bf16": MATH_BF16, "fp32": MATH_FP32}[args.math])
    eager_step = TrainStep(model, lr=1e-3, communicator=comm)
    # launch-bound regime (cfgA): replay the whole step as one CUDA graph
    graphed = args.graph if args.graph is not None else (args.workload == "cfgA" and world == 1)
    if graphed:
        from gnn_link_prediction_b200.train import GraphedTrainStep
        step = GraphedTrainStep(eager_step)
    else:
        step = eager_step
    dev_batches = [copy_batch_to_device(h) for h in host]
    torch.cuda.synchronize()
    # one instrumented eager step: counts the kernels a step launches (graph replays cannot be
    # instrumented per kernel) and gives the eager per-kernel breakdown
    eager_step(dev_batches[0])   # un-instrumented first call: lazy init, allocator warm-up
    torch.cuda.synchronize()
    probe = KernelTimer()
    ops.TIMER = probe
    eager_step(dev_batches[1 % n_host])
    ops.TIMER = None
    probe_kernels = probe.summary()
    kernels_per_step = sum(k.get("kernels", k["launches"]) for k in probe_kernels.values())   # wrapper-side estimate
    # exact count: one eager step under the CUPTI-based profiler, kernels of libhgin only (namespace hgin::)
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eager_step(dev_batches[0])
            torch.cuda.synchronize()
        counted = sum(1 for e in prof.events() if "hgin::" in e.name)
        if counted > 0:
            kernels_per_step = counted
    except Exception:
        pass

    small = args.workload == "cfgA"   # working set << L2: flush between timed steps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if small else None

    def barrier():
        comm.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ("value") ----------------------------------------------------------
    packed_host = packed_dev = None
    if graphed:
        # graph mode: each batch is ONE pinned buffer (edge lists padded with (-1,-1) slots)
        from gnn_link_prediction_b200.data import PackedBatch, pack_batch
        packed_host = [pack_batch(h) for h in host]
        packed_dev = [PackedBatch(p.buffer.cuda(), p.layout, p.num_graphs) for p in packed_host]
        h2d_bytes = packed_host[0].nbytes()

    def resident_step(i):
        step(packed_dev[i % n_host] if graphed else dev_batches[i % n_host])

    for i in range(args.warmup):
        resident_step(i)
    timer = KernelTimer()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ops.TIMER = None if graphed else timer   # per-kernel events cannot be recorded inside a graph replay
    t_wall = time.perf_counter()
    per_step = timed_steps(resident_step, args.steps, flush)
    barrier()
    t_wall = time.perf_counter() - t_wall
    ops.TIMER = None
    clocks = sampler.result()
    kernels = timer.summary() if not graphed else probe_kernels
    timed_steps_for_kernels = args.steps if not graphed else 1
    resident_ms = sum(per_step)

    # ---- end-to-end arm ("e2e"): pinned host batch -> H2D -> step -> D2H loss, every step ---------
    # Through the public API a user would call: DevicePrefetcher stages batch i+1 on a copy stream
    # while step i runs; every step's copy and its loss read-back are inside the timed region (the
    # first copy is exposed, the rest overlap compute).
    from gnn_link_prediction_b200.data import DevicePrefetcher
    from gnn_link_prediction_b200.train import LossReadback
    losses = []
    e2e_host = host
    if not graphed and args.stage == "packed":
        from gnn_link_prediction_b200.data import pack_batch
        e2e_host = [pack_batch(h) for h in host]       # loader-side work, like the collate itself
        h2d_bytes = e2e_host[0].nbytes()
    prefetcher = DevicePrefetcher(())

    def prefetcher_over(source):      # one prefetcher (one device ring) for the whole run
        prefetcher.source = source
        return prefetcher

    def e2e_run(n_steps):
        # every step's loss is read back to the host inside the timed region; with the default
        # `deferred` read-back (train.LossReadback) the host collects step i's loss after launching
        # step i+1, `sync` is the reference's blocking `.item()` (train.py:50)
        reader = LossReadback() if args.readback == "deferred" else None

        def collect(loss):
            if reader is None:
                losses.append(loss.cpu())
            else:
                done = reader.push(loss)
                if done is not None:
                    losses.append(done)

        if graphed:   # one H2D copy of the packed pinned batch into the graph's static buffer, replay, D2H
            for i in range(n_steps):
                collect(step(packed_host[i % n_host]))
        else:         # packed pinned batch -> one DMA into the prefetcher's device ring, overlapping the previous step
            for dev in prefetcher_over(e2e_host[i % n_host] for i in range(n_steps)):
                collect(step(dev))
        if reader is not None:
            losses.append(reader.flush())

    e2e_run(2)
    barrier()
    if flush is not None:
        flush.zero_()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    e2e_run(args.steps)
    eb.record()
    barrier()
    e2e_ms = ea.elapsed_time(eb)

    # ---- device-resident dataset arm: the samples live in HBM (arena.DeviceDataset), every step assembles a
    # FRESH random batch on the GPU from its sample ids (H2D = the id list), then runs the same step.
    dd_ms, dd_collate_ms, dd_bytes = 0.0, 0.0, 0
    if not graphed and args.collate == "csr":
        from gnn_link_prediction_b200.arena import DeviceDataset, SampleArena
        dds = DeviceDataset(SampleArena.from_samples(ALL_SAMPLES, keep_coo=False))
        gen = torch.Generator().manual_seed(1997 + rank)
        id_lists = [torch.randint(0, len(dds), (graphs,), generator=gen).numpy() for _ in range(args.steps + 2)]
        dd_bytes = dds.h2d_bytes(graphs)

        def dd_run(lists):
            reader = LossReadback()
            for ids in lists:
                reader.push(step(dds.collate(ids)))
            reader.flush()

        dd_run(id_lists[:2])
        barrier()
        da, db_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        da.record()
        dd_run(id_lists[2:])
        db_.record()
        barrier()
        dd_ms = da.elapsed_time(db_)
        ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ca.record()
        for ids in id_lists[2:]:
            dds.collate(ids)
        cb.record()
        torch.cuda.synchronize()
        dd_collate_ms = ca.elapsed_time(cb) / args.steps

    # ---- host-streaming arm: the samples stay in HOST memory (arena.SampleArena); a background thread assembles a
    # FRESH shuffled batch per step with the native host collate into a ring of pinned buffers (arena.HostLoader),
    # DevicePrefetcher copies it (one DMA), the step runs, the loss is read back.  Collate, H2D and D2H are all
    # inside the timed region; this arm is bound by the host's memory bandwidth, not by the GPU.
    hs_ms, hs_steps = 0.0, 0
    if not graphed and args.collate == "csr":
        from gnn_link_prediction_b200.arena import HostLoader
        per_epoch = len(dds.arena) // graphs
        hl = HostLoader(dds.arena, batch_size=graphs, shuffle=True, generator=torch.Generator().manual_seed(7 + rank),
                        ring=4, epochs=(args.steps + 2 + per_epoch - 1) // per_epoch,
                        num_threads=max(1, min(8, (os.cpu_count() or 8) // world)))
        reader = LossReadback()
        ha, hb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k, dev in enumerate(prefetcher_over(hl)):
            if k == 2:                      # two warm-up batches (pinned ring allocation), then the timed ones
                barrier()
                ha.record()
            reader.push(step(dev))
            if k >= 2:
                hs_steps += 1
            if hs_steps == args.steps:
                break
        reader.flush()
        hb.record()
        barrier()
        hs_ms = ha.elapsed_time(hb)

    # ---- strong-scaling arm: the SAME global batch (w["batch"] topologies) split over the
    # ranks, batch resident in HBM; reported beside the weak-scaling `value`.
    strong_ms, strong_graphs = 0.0, 0
    if world > 1 and not graphed and w["batch"] % world == 0:
        from gnn_link_prediction_b200.data import Batch, CONV_EDGE_TYPES
        share = w["batch"] // world
        strong_graphs = share * world
        sb = copy_batch_to_device(Batch.from_data_list(ALL_SAMPLES[:share], index_dtype=torch.int32,
                                                       edge_types=CONV_EDGE_TYPES, batch_vector=False, csr=True,
                                                       keep_coo=False))
        for _ in range(args.warmup):
            step(sb)
        barrier()
        strong_ms = sum(timed_steps(lambda i: step(sb), args.steps, flush))
        barrier()

    # max over ranks (device time)
    t = torch.tensor([resident_ms, e2e_ms, dd_ms, hs_ms, strong_ms], dtype=torch.float64, device="cuda")
    comm.all_reduce_max_(t)
    resident_ms, e2e_ms, dd_ms, hs_ms, strong_ms = (float(v) for v in t.tolist())
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    hbm_peak, tf_peak, basis = measured_peaks()
    agg = kernels.get("gin_combine", {"ms": 0.0, "launches": 0, "alg_bytes": 0, "compulsory_bytes": 0})
    total_kernel_ms = sum(k["ms"] for k in kernels.values()) or 1.0
    roofline = None
    if agg["launches"]:
        # block-diagonal batches: every source row is fetched from DRAM once and re-read from L2/L1,
        # so the COMPULSORY byte count is the one that can be held against the HBM peak (it matches
        # the ncu DRAM traffic within 2%, profiles/); the algorithmic (per-edge) figure is beside it.
        achieved = agg["compulsory_bytes"] / (agg["ms"] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload, {}).get("gin_combine_dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": "hgin_gin_combine", "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_basis": basis,
                    "bytes": "compulsory bytes of SURVEY 8(d) (block-diagonal batch); traffic = ncu dram bytes per launch",
                    "launches": agg["launches"], "avg_launch_ms": agg["ms"] / agg["launches"],
                    "share_of_step": agg["ms"] / total_kernel_ms,
                    "achieved_algorithmic_GBs": agg["alg_bytes"] / (agg["ms"] * 1e-3) / 1e9}
    breakdown = {name: {"ms_per_step": k["ms"] / timed_steps_for_kernels,
                        "launches_per_step": k["launches"] / timed_steps_for_kernels,
                        **({"TFLOPs": k["flops"] / (k["ms"] * 1e-3) / 1e12} if "flops" in k and k["ms"] > 0 else {}),
                        **({"GBs": k["bytes"] / (k["ms"] * 1e-3) / 1e9} if "bytes" in k and k["ms"] > 0 else {})}
                 for name, k in kernels.items()}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sample, cpu_steps = CPU_SAMPLE[args.workload]
        r = cpu_reference_run(w, sample, cpu_steps, 1, threads=host_cores())
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    launches = kernels_per_step * args.steps
    value = graphs * world * args.steps / (resident_ms * 1e-3)
    e2e_pre = {"value": graphs * world * args.steps / (e2e_ms * 1e-3), "unit": "graphs/s",
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms / args.steps,
               "collate": "outside the timed region (two pre-collated, pre-packed pinned batches cycled)",
               "staging": "one packed pinned buffer per batch -> ping-pong static buffers of two CUDA graphs, copied on a copy stream" if graphed
               else ("one packed pinned buffer per batch -> DevicePrefetcher ring" if args.stage == "packed"
                     else "one H2D copy per tensor (DevicePrefetcher)"),
               "readback": "every step, collected one step later (train.LossReadback)" if args.readback == "deferred"
               else "every step, blocking"}
    # the host-streaming arm: a FRESH shuffled batch is collated from host-resident samples INSIDE the timed region
    # every step (native host collate on a background thread into a pinned ring), copied with one DMA, trained on, and
    # its loss read back.  Bound by the HOST's memory system, not by the GPU: per rank and step the collate reads and
    # writes 230 MB and the DMA reads them again, so 8 ranks on one host ask for ~350 GB/s of host memory traffic.
    e2e_host = None if hs_steps == 0 else {
        "value": graphs * world * hs_steps / (hs_ms * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": h2d_bytes,
        "d2h_bytes_per_step": 8, "ms_per_step": hs_ms / hs_steps,
        "collate": "inside the timed region (arena.HostLoader: native host collate of a fresh shuffled batch per step)",
        "staging": "one packed pinned buffer per batch -> DevicePrefetcher ring (one DMA per step)",
        "readback": "every step, collected one step later (train.LossReadback)"}
    # the parsed end-to-end number = the production data path: the dataset is resident in HBM
    # (arena.DeviceDataset; 225 KB per sample, so 180 GB hold ~800 k samples), every step copies that step's sample ids
    # from pinned host memory, assembles a FRESH random batch on the GPU (hgin_collate_*), trains on it and reads the
    # loss back — all inside the timed region.  (cfgA under a CUDA graph has no loader arm: pre-collated.)
    e2e_main = e2e_pre if dd_ms == 0.0 else {
        "value": graphs * world * args.steps / (dd_ms * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": dd_bytes,
        "d2h_bytes_per_step": 8, "ms_per_step": dd_ms / args.steps, "collate_ms_per_step": dd_collate_ms,
        "collate": "inside the timed region, on the GPU (arena.DeviceDataset: fresh random batch per step from the sample ids)",
        "staging": "dataset resident in HBM (uploaded once, outside the timed region); per step: H2D of the sample ids "
                   "from pinned host memory",
        "readback": "every step, collected one step later (train.LossReadback)"}
    line = {
        "metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": resident_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32 (dense layers; f32 aggregation and accumulation)",
                                     "bf16": "bf16 (activations / gradients stored as bf16, tcgen05 bf16 GEMMs; f32 accumulation, "
                                             "aggregation adds, loss, optimizer)"}[args.math], "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "emb": w["emb"], "layers": w["layers"],
                   "graphs_per_gpu_per_step": graphs, "edges_per_gpu_per_step": edges,
                   "l2": "flushed between timed steps" if small else "inputs+activations larger than L2",
                   "collate": args.collate,
                   "parallelism": f"dp{world} (samples sharded, NCCL sum-allreduce of one flat grad bucket)"},
        "edges_per_s": edges * world * args.steps / (resident_ms * 1e-3),
        "e2e": e2e_main,
        "e2e_host_stream": e2e_host,
        "e2e_precollated": e2e_pre if e2e_main is not e2e_pre else None,
        "strong_scaling": None if strong_ms == 0.0 else {
            "value": strong_graphs * args.steps / (strong_ms * 1e-3), "unit": "graphs/s",
            "ms_per_step": strong_ms / args.steps, "global_batch": strong_graphs, "graphs_per_gpu_per_step": strong_graphs // world,
            "what": "strong scaling: the 1-GPU global batch split over the ranks (batch resident in HBM), beside the "
                    "weak-scaling `value`"},
        "gpu_launches": launches, "kernels_per_step": kernels_per_step,
        "execution": "one CUDA graph replay per step (GraphedTrainStep)" if graphed else "eager launches",
        "wall_s_resident": t_wall, "peak_hbm_gb": torch.cuda.max_memory_allocated() / 1e9,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "kernels": breakdown,
        "loss_first_last": [float(losses[0][0]), float(losses[-1][0])],
    }
    if args.workload == "cfgC" and world == 1 and not args.no_extra:
        # BASELINE configs[1] (config.json defaults, launch-bound) beside the default workload: same script, own process
        import subprocess
        try:
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "cfgA", "--no-cpu-baseline",
                                "--no-extra", "--steps", str(max(args.steps, 20)), "--warmup", str(args.warmup)],
                               capture_output=True, text=True, timeout=300)
            a = json.loads(r.stdout.strip().splitlines()[-1])
            line["other_workloads"] = {"cfgA": {"desc": a["config"]["desc"], "value": a["value"], "unit": a["unit"],
                                                "ms_per_step": a["ms_per_step"], "e2e": a["e2e"], "execution": a["execution"],
                                                "kernels_per_step": a["kernels_per_step"], "dtype": a["dtype"]}}
        except Exception as exc:   # the default line must not depend on the extra run
            line["other_workloads"] = {"cfgA": {"error": repr(exc)[:200]}}
        # the same Cfg-C workload in the bf16 storage mode (BASELINE configs[2] "bf16 MLP GEMMs"): a second line BESIDE
        # the tf32 one, with its own roofline, measured by the same script in its own process
        if args.math != "bf16":
            try:
                torch.cuda.empty_cache()
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "cfgC", "--math", "bf16",
                                    "--no-cpu-baseline", "--no-extra", "--steps", str(args.steps), "--warmup", str(args.warmup)],
                                   capture_output=True, text=True, timeout=300)
                b = json.loads(r.stdout.strip().splitlines()[-1])
                line["other_workloads"]["cfgC_bf16"] = {k: b[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e", "e2e_host_stream", "roofline",
                                                                          "edges_per_s", "kernels_per_step", "peak_hbm_gb")}
            except Exception as exc:
                line["other_workloads"]["cfgC_bf16"] = {"error": repr(exc)[:200]}
        # and the same tf32 step captured once and replayed as ONE CUDA graph (train.GraphedTrainStep; single GPU only —
        # NCCL inside the capture blocked twice, DESIGN.md §7 — so the headline stays on eager launches, which scale out)
        if not graphed:
            try:
                torch.cuda.empty_cache()
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "cfgC", "--math", args.math, "--graph",
                                    "--no-cpu-baseline", "--no-extra", "--steps", str(args.steps), "--warmup", str(args.warmup)],
                                   capture_output=True, text=True, timeout=300)
                g = json.loads(r.stdout.strip().splitlines()[-1])
                line["other_workloads"]["cfgC_graph_replay"] = {k: g[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e",
                                                                                  "execution", "kernels_per_step")}
            except Exception as exc:
                line["other_workloads"]["cfgC_graph_replay"] = {"error": repr(exc)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
