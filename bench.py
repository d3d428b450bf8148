#!/usr/bin/env python
"""Benchmark of the per-read simulation hot path (BASELINE.json metric: simulated bases/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch_reads B]

Workload (BASELINE config 2): genome mode, human_NA12878_DNA_FAB49712_guppy error/length model with the
dorado_v3.2.1 base-quality table (the shipped guppy model has none), FASTQ, on a 3.09 Gb synthetic
reference (24 chromosomes with hg38 lengths, i.i.d. ACGT).  The whole job is 10M reads; ONE STEP is one batch of
``--batch_reads`` reads (aligned + unaligned in the model's 8.85:1 ratio) = the unit the job is made of, so
bases/sec over K steps is the job's throughput.  Under torchrun every rank simulates its own batch per step
(weak scaling; read ids are disjoint shards) after ONE NCCL broadcast of the reference at init.

Printed JSON line: see the task contract.  ``value`` = bases / device time of the kernels (outputs stay in HBM);
``e2e`` = the same through ns_simulate + ns_fetch into pinned host buffers (D2H inside the timed region);
``roofline`` = emit kernel, 3 algorithmic bytes per base (1 reference byte read + 1 base + 1 quality written);
``cpu_baseline`` = the oracle port (pure Python, like the reference) on a bounded sample with all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
CHROM_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]
MODEL = os.path.join(ROOT, "nanosim_b200", "data", "guppy_fab49712_plusq.npz")
ALGO_BYTES_PER_BASE = 3.0
# dram__bytes_read.sum + dram__bytes_write.sum of ONE emit_kernel<FASTQ> launch of this workload (235536 aligned reads,
# 2.04 Gbases): 3.226 GB + 4.401 GB, from `ncu --set full -k regex:emit_kernel -c 1 python bench.py --steps 1 --warmup 1
# --depth 1 --no_cpu_baseline` (profiles/r1_emit_ncu_summary.txt) = 3.73 B/base against 3 B/base algorithmic.
TRAFFIC_BYTES_PER_LAUNCH = 7.627e9


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from the warm-up on (nvidia-smi itself needs ~0.1 s to
    start, longer than a short timed region); samples are time-stamped and only the ones inside a timed region count."""

    def __init__(self, device):
        self.rows = []
        self.device = device
        self.proc = None
        self.windows = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1, label):
        self.windows.append((t0, t1, label))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if any(a <= t <= b + 0.06 for a, b, _ in self.windows)]
        sm = [float(r[0]) for r in inside if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for _, r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = sorted({names[i] for r in inside if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                "windows": "samples inside the timed regions: " + ", ".join("%s %.0f ms" % (lab, 1e3 * (b - a)) for a, b, lab in self.windows)}


def synth_reference_gpu(device, scale=1.0):
    """24 chromosomes, i.i.d. uniform ACGT, generated on the GPU (torch is plumbing here)."""
    import torch

    lengths = [max(1000, int(x * scale)) for x in HG38_LENGTHS]
    total = sum(lengths)
    g = torch.Generator(device="cuda:%d" % device)
    g.manual_seed(1)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda:%d" % device)
    out = torch.empty(total, dtype=torch.uint8, device="cuda:%d" % device)
    step = 1 << 28
    for s in range(0, total, step):
        e = min(total, s + step)
        idx = torch.randint(0, 4, (e - s,), generator=g, device="cuda:%d" % device, dtype=torch.uint8)
        out[s:e] = lut[idx.long()]
        del idx
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    return out, offsets


def cpu_oracle_sample(ref_strs, n_reads, n_procs, fastq=True):
    """Times the oracle port (pure Python, same algorithm and data structures as the reference) on ``n_reads`` reads
    split over ``n_procs`` forked workers, like ``simulator.py -t``.  Returns (bases, seconds)."""
    import multiprocessing as mp

    from conftest import oracle_model
    from nanosim_b200.model import CompiledModel
    import nanosim_oracle as no

    cm = CompiledModel.load(MODEL)
    tmp = tempfile.mkdtemp(prefix="bench_oracle_")
    m = oracle_model(cm, tmp, fastq=fastq)
    oref = no.OracleReference(ref_strs)
    n_al, n_un = m.split_counts(n_reads)
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def work(i, na, nu):
        import random
        random.seed(1000 + i)
        np.random.seed(1000 + i)
        s1, s2 = no.ReadSink(), no.ReadSink()
        no.simulation_aligned_genome(oref, m, s1, "linear", 50, oref.max_chrom, None, None, None, fastq, na, False, False)
        if nu:
            no.simulation_unaligned(oref, m, s2, "linear", 50, oref.max_chrom, None, None, fastq, nu)
        txt = no.format_records(s1.records + s2.records, fastq)      # the reference also formats and writes records
        q.put((sum(len(r[1]) for r in s1.records + s2.records), len(txt)))

    t0 = time.time()
    procs = []
    for i in range(n_procs):
        na = n_al // n_procs + (n_al % n_procs if i == n_procs - 1 else 0)
        nu = n_un // n_procs + (n_un % n_procs if i == n_procs - 1 else 0)
        p = ctx.Process(target=work, args=(i, na, nu))
        p.start()
        procs.append(p)
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    dt = time.time() - t0
    return sum(r[0] for r in res), dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch_reads", type=int, default=262144)
    ap.add_argument("--ref_scale", type=float, default=1.0, help="scale the 3.09 Gb reference (tests only)")
    ap.add_argument("--depth", type=int, default=4, help="overlapped contexts per GPU")
    ap.add_argument("--timeline", default=None, help="write the per-batch phase intervals of the timed steps to this file")
    ap.add_argument("--cpu_reads", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = "genome FASTQ, guppy FAB49712 model + dorado_v3.2.1 quality table, 3.09 Gb synthetic hg38-sized reference " \
               "(24 chr, i.i.d. ACGT), 10M-read job, %d reads per step" % args.batch_reads

    import torch

    if args.impl == "reference":
        # the reference's CPU implementation of the path == the oracle port (Python, multiprocessing over all cores)
        if rank != 0:
            return
        ref_t, offsets = synth_reference_gpu(local, args.ref_scale) if torch.cuda.is_available() else (None, None)
        if ref_t is None:
            rng = np.random.default_rng(1)
            lengths = [max(1000, int(x * args.ref_scale)) for x in HG38_LENGTHS]
            host = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)] for n in lengths]
            ref_strs = [(nm, a.tobytes().decode()) for nm, a in zip(CHROM_NAMES, host)]
        else:
            host = ref_t.cpu().numpy()
            ref_strs = [(nm, host[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()) for i, nm in enumerate(CHROM_NAMES)]
            del ref_t
        per_step = args.cpu_reads or 40 * cores
        vals = []
        for s in range(args.warmup + args.steps):
            bases, dt = cpu_oracle_sample(ref_strs, per_step, cores)
            if s >= args.warmup:
                vals.append((bases, dt))
        tb, tt = sum(v[0] for v in vals), sum(v[1] for v in vals)
        v = tb / tt
        print(json.dumps({"impl": "reference", "metric": "simulated_bases_per_sec", "value": v, "unit": "bases/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * tt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": workload},
                          "cpu_baseline": {"value": v, "unit": "bases/s", "cores": cores, "kind": "port",
                                           "sample": "%d reads per step, %d forked workers, pure-Python oracle port of simulator.py "
                                                     "(same algorithm as the reference, incl. record formatting)" % (per_step, cores)},
                          "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from nanosim_b200 import _lib as L
    from nanosim_b200.engine import Engine
    from nanosim_b200.model import CompiledModel, DeviceTables

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))

    # ---- init (not timed): reference generated on rank 0, ONE broadcast over NCCL, model tables to HBM
    if rank == 0:
        ref_t, offsets = synth_reference_gpu(local, args.ref_scale)
    else:
        lengths = [max(1000, int(x * args.ref_scale)) for x in HG38_LENGTHS]
        offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
        ref_t = torch.empty(int(offsets[-1]), dtype=torch.uint8, device="cuda:%d" % local)
    if world > 1:
        dist.broadcast(ref_t, src=0)
    torch.cuda.synchronize()
    cm = CompiledModel.load(MODEL)
    tables = DeviceTables(cm, fastq=True)
    eng = Engine(device=local, seed=20260924)
    eng.set_reference_ptr(ref_t.data_ptr(), int(offsets[-1]), offsets)
    host_ref = ref_t.cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None     # CPU baseline: N=1 only
    del ref_t
    torch.cuda.empty_cache()
    eng.set_model(tables)
    eng.configure(fastq=True, min_len=50, max_len=int(np.diff(offsets.astype(np.int64)).max()))

    n_al, n_un = tables.split_counts(args.batch_reads)
    total_steps = args.warmup + args.steps
    from nanosim_b200.pipeline import BatchPipeline

    def jobs_for(steps):
        """One step = one batch of the job on this rank: its aligned reads, then its unaligned reads."""
        out = []
        for step in steps:
            if n_al:
                out.append((L.NS_KIND_ALIGNED, (step * world + rank) * n_al, n_al))
            if n_un:
                out.append((L.NS_KIND_UNALIGNED, (step * world + rank) * n_un, n_un))
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def row(info):
        return (info.total_bases, info.seq_bytes, info.n_reads, info.n_pieces, info.n_launches, info.ms_total,
                info.ms_plan, info.ms_scan, info.ms_script, info.ms_emit, info.ms_setup, info.t_begin_ms, info.t_end_ms)

    # ---- kernel-only arm: outputs stay in HBM.  `depth` contexts (ns_clone) share the reference; the latency-bound tails
    #      of one batch's plan / unaligned kernels overlap the emit kernel of another.  Every step simulates new read ids;
    #      a batch's working set (>2 GB written + a 3 GB reference sampled at random) is far larger than the 126 MB L2.
    pipe = BatchPipeline(eng, depth=args.depth, fetch=False)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    pipe.warm(jobs_for(range(1)))                   # every context sizes its buffers once (untimed)
    pipe.run(jobs_for(range(args.warmup)))
    barrier()
    t0 = time.perf_counter()
    rows = [row(i) for i in pipe.run(jobs_for(range(args.warmup, total_steps)))]
    barrier()
    wall = time.perf_counter() - t0
    clocks.window(t0, t0 + wall, "kernel-only arm")
    pipe.close()
    if args.timeline and rank == 0:
        with open(args.timeline, "w") as f:       # phases are back to back on a context's stream: begin + cumulative durations
            f.write("reads\tbegin\tsetup_end\tplan_end\tscan_end\tscript_end\temit_end\n")
            t00 = min(r[11] for r in rows)
            for r in sorted(rows, key=lambda r: r[11]):
                t = r[11] - t00
                cells = [t, t + r[10], t + r[10] + r[6], t + r[10] + r[6] + r[7], t + r[10] + r[6] + r[7] + r[8], r[12] - t00]
                f.write("%d\t%s\n" % (r[2], "\t".join("%.2f" % c for c in cells)))
    bases = sum(r[0] for r in rows)
    dev_ms = max(r[12] for r in rows) - min(r[11] for r in rows)     # device timeline: first batch start -> last batch end
    emit_ms = sum(r[9] for r in rows)
    launches = sum(r[4] for r in rows)
    n_reads_done = sum(r[2] for r in rows)
    stat = torch.tensor([bases, dev_ms, wall * 1e3, n_reads_done], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        mx = stat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_bases, t_ms, total_reads = float(sm[0]), float(mx[1]), float(sm[3])
    else:
        total_bases, t_ms, total_reads = bases, dev_ms, n_reads_done
    value = total_bases / (t_ms * 1e-3)

    # ---- roofline leg: the dominant kernel timed ALONE.  With several overlapped contexts the CUDA events around a
    #      launch also span the other contexts' kernels sharing the SMs, so the per-kernel durations of the arm above
    #      over-state every kernel; here the same batches run through one context (fresh read ids, same sizes).
    n_roof = max(1, min(args.steps, 3))
    pipe1 = BatchPipeline(eng, depth=1, fetch=False)
    pipe1.run(jobs_for(range(total_steps, total_steps + 1)))
    barrier()
    rows1 = [row(i) for i in pipe1.run(jobs_for(range(total_steps + 1, total_steps + 1 + n_roof)))]
    barrier()
    pipe1.close()
    al1 = [r for r in rows1 if r[2] == n_al]                         # aligned batches -> emit_kernel<FASTQ> launches
    emit_alone_ms = sum(r[9] for r in al1) / max(len(al1), 1)
    emit_alone_bases = sum(r[0] for r in al1) / max(len(al1), 1)
    plan_alone_ms = sum(r[6] for r in al1) / max(len(al1), 1)
    total_steps += 1 + n_roof

    # ---- end-to-end arm: the public API (BatchPipeline): ns_simulate + ns_fetch into pinned host buffers every batch
    pipe_e = BatchPipeline(eng, depth=args.depth, fetch=True)
    base_step = total_steps                          # fresh read ids
    e_warm = max(3, args.depth + 1)                  # every context's pinned buffers must have seen an aligned batch
    pipe_e.warm(jobs_for(range(base_step, base_step + 1)))
    pipe_e.run(jobs_for(range(base_step, base_step + e_warm)))
    barrier()
    t0 = time.perf_counter()
    rows_e = [row(i) for i in pipe_e.run(jobs_for(range(base_step + e_warm, base_step + e_warm + args.steps)))]
    barrier()
    wall_e = time.perf_counter() - t0
    clocks.window(t0, t0 + wall_e, "end-to-end arm")
    clk = clocks.stop() if rank == 0 else None
    pipe_e.close()
    bases_e = sum(r[0] for r in rows_e)
    # bytes that cross PCIe: qualities as ASCII, bases as 2 bits (packed on the device, expanded by host threads inside
    # ns_fetch unless NANOSIM_B200_UNPACK_THREADS=0) + read / piece metadata
    env_t = os.environ.get("NANOSIM_B200_UNPACK_THREADS", "")          # same rule as unpack_threads() in nanosim_api.cu
    packed = int(env_t) > 0 if env_t else (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))) >= 48
    d2h = sum((1.25 if packed and r[1] >= (1 << 20) else 2.0) * r[1] + 32 * r[2] + 64 * r[3] for r in rows_e) / max(args.steps, 1)
    stat = torch.tensor([bases_e, wall_e], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        mx = stat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_value = float(sm[0]) / float(mx[1])
    else:
        e2e_value = bases_e / wall_e

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_BASE * emit_alone_bases / (emit_alone_ms * 1e-3) / 1e9
    line = {
        "metric": "simulated_bases_per_sec", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "reads_per_step_per_gpu": args.batch_reads, "aligned_per_step": n_al,
                   "unaligned_per_step": n_un, "l2": "inputs larger than L2 (3.09 GB reference sampled at random, >2 GB written per step)",
                   "contexts_per_gpu": args.depth,
                   "timing": "device timeline (CUDA events vs a common base event): first batch start to last batch end of the K "
                             "timed steps, %d overlapped contexts per GPU, max over ranks" % args.depth},
        "reads_per_sec": total_reads / (t_ms * 1e-3),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": int(d2h),
                "note": "reference + model are resident in HBM (uploaded once at init); per-step input is the read-id range; the "
                        "timed region ends with ASCII bases + qualities + metadata in pinned host buffers" +
                        (" (bases cross PCIe as 2 bits and are expanded by host threads inside ns_fetch)" if packed else "")},
        "gpu_launches": int(launches),
        "phase_ms_per_step": {"plan": sum(r[6] for r in rows) / args.steps, "scan": sum(r[7] for r in rows) / args.steps,
                              "script": sum(r[8] for r in rows) / args.steps, "emit": emit_ms / args.steps,
                              "setup": sum(r[10] for r in rows) / args.steps,
                              "note": "sums of per-batch CUDA-event durations; batches of the overlapped contexts share the GPU, so these add up to more than ms_per_step"},
        "wall_ms_per_step": 1e3 * wall / max(args.steps, 1),
        "roofline": {"bound": "hbm", "kernel": "emit_kernel<FASTQ>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH, "peak_source": peak_src,
                     "algorithmic_bytes_per_base": ALGO_BYTES_PER_BASE,
                     "bases_per_launch": emit_alone_bases, "ms_per_launch": emit_alone_ms,
                     "plan_kernel_ms_per_launch": plan_alone_ms,
                     "measured": "CUDA events on the launching stream around emit_kernel<FASTQ>, %d aligned batches of %d reads run "
                                 "through ONE context after the timed region (kernels of overlapped contexts share SMs, which "
                                 "stretches every per-launch duration)" % (len(al1), n_al),
                     "whole_path_frac": ALGO_BYTES_PER_BASE * total_bases / (t_ms * 1e-3) / 1e9 / peak / max(world, 1)},
    }
    if not args.no_cpu_baseline and host_ref is not None:
        ref_strs = [(nm, host_ref[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()) for i, nm in enumerate(CHROM_NAMES)]
        n_cpu = args.cpu_reads or 40 * cores
        cb, ct = cpu_oracle_sample(ref_strs, n_cpu, cores)
        line["cpu_baseline"] = {"value": cb / ct, "unit": "bases/s", "cores": cores, "kind": "port",
                                "sample": "%d reads of the same workload, %d forked workers, pure-Python oracle port of simulator.py "
                                          "(record formatting included, no file I/O); %.1f s" % (n_cpu, cores, ct)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
