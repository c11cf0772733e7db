#!/usr/bin/env python
"""bench/msa_bench.py — MultiStateAligner11ts microbenchmark (BASELINE.json configs[2], SURVEY.md §8d workload G4).

One "step" = one pass of the hot path (fillLimited rule -> fillLimitedX/fillUnlimited -> score2 -> traceback2) over one
batch of synthetic (read, candidate window) tasks.  Metric: DP GCUPS = the reference's own cell counter
(iterationsLimited+iterationsUnlimited, jni/MultiStateAligner11tsJNI.c:138,471) summed over the batch / seconds.

  python bench/msa_bench.py [--gpus N] [--steps K] [--warmup W] [--tasks T] [--impl reference] [--bandwidth B --ratio R]

bench.py (repo root) embeds this measurement as `msa` next to the mapped-reads/s headline.

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the host-buffer
plug-in call (H2D of tasks+reads and D2H of results inside the timed region).  `--impl reference` times the
reference's own C (oracle/_ref, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bbmap_b200 import workloads as wl  # noqa: E402

GENOME_LEN = 4_600_000      # "E. coli-sized" resident reference the windows point into
METRIC = "MSA fill GCUPS (MultiStateAligner11ts fillLimited+score+traceback, reference cell count / s)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tasks", type=int, default=1_600_000, help="alignments per step per GPU")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--bandwidth", type=int, default=0)
    ap.add_argument("--ratio", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lengths", default="100,150,250")
    ap.add_argument("--no-narrow", action="store_true", help="disable the thread-per-alignment narrow kernel (A/B)")
    ap.add_argument("--no-strip", action="store_true", help="route limited fills through the register-tiled kernel instead of the strip kernel (A/B)")
    ap.add_argument("--strip-budget-mb", type=int, default=0)
    ap.add_argument("--narrow-slack", type=int, default=-1, help="narrow kernel only for alignments with maxQ-minScore <= this (points)")
    ap.add_argument("--strip-buckets", type=int, default=-1, help="work buckets (of 4096 cells) routed to the strip kernel; larger alignments use the tiled kernel")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-stage timings (ingest/seed/index/search/scoreNoIndels) of bench/stages.py")
    ap.add_argument("--stage-pairs", type=int, default=200_000)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(reads, genome, tasks, moff, bw, ratio, budget_s=12.0, threads=None):
    """Times the reference's own C (oracle/_ref) — or the port if it is absent — on a bounded sample: whole passes over
    (a prefix of) the step batch until ~budget_s seconds of wall time have been spent."""
    from oracle import oracle as orc
    o = orc.get()
    kind = "reference" if o.has_reference else "port"
    threads = threads or (os.cpu_count() or 1)
    n = int(min(len(tasks), 100_000))
    cells = 0; secs = 0.0; passes = 0
    while secs < budget_s and passes < 64:
        t0 = time.perf_counter()
        _, _, c = o.run_batch(reads, genome, tasks[:n], match_off=moff[:n + 1], bandwidth=bw, ratio=ratio, kind=kind, threads=threads)
        secs += time.perf_counter() - t0
        cells += c; passes += 1
    return {"value": cells / secs / 1e9, "unit": "GCUPS", "cores": threads, "kind": kind,
            "sample": "%d passes over the first %d alignments of the step batch, %.1f s, %d threads, one private packed matrix per thread"
                      % (passes, n, secs, threads),
            "seconds": secs, "cells": cells, "tasks": n * passes}


def reference_line(args):
    """The reference's own C (oracle/_ref) on a bounded sample; returns the JSON-able dict."""
    return measure(args, 0, 1, 0, reference=True)


def measure(args, rank, world, local, reference=False):
    """One measurement; the caller has initialised torch.distributed (gloo) when world > 1.  Returns the line (rank 0) or None."""
    if reference:
        args.impl = "reference"
    lengths = tuple(int(x) for x in args.lengths.split(","))
    config = {"workload": "configs[2] MSA11ts microbenchmark G4: %d alignments/step/GPU, read length {%s}, window = locus +-4, "
                          "70%% ~1%% subs / 20%% 1-40bp indel / 10%% unrelated, minScore=max(scoreNoIndels, 0.56*maxQ-258), "
                          "fillLimited rule + score2 + traceback2; resident %d bp reference" % (args.tasks, args.lengths, GENOME_LEN),
              "tasks_per_step_per_gpu": args.tasks, "bandwidth": args.bandwidth, "bandwidthRatio": args.ratio,
              "l2": "inputs larger than L2 (tasks+reads+outs+match > 200 MB per step)"}

    genome = wl.random_genome(GENOME_LEN, seed=1)
    # the reference arm times a bounded sample (the first 20 k alignments of the step batch): generate only the first block
    ngen = args.tasks if args.impl != "reference" else min(args.tasks, 100_000)
    reads, tasks = wl.make_msa_tasks(genome, ngen, seed=2 + rank, lengths=lengths, flags=wl.TF_SCORE | wl.TF_TRACEBACK)
    moff = wl.match_offsets(tasks)

    if args.impl == "reference":
        if rank != 0:
            return None
        # each step = a bounded sample sized so steps+warmup finish within minutes
        from oracle import oracle as orc
        o = orc.get()
        kind = "reference" if o.has_reference else "port"
        threads = os.cpu_count() or 1
        n = min(len(tasks), 20000)
        times, cells = [], 0
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _, _, c = o.run_batch(reads, genome, tasks[:n], match_off=moff[:n + 1], bandwidth=args.bandwidth, ratio=args.ratio,
                                  kind=kind, threads=threads)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt); cells += c
        total = sum(times)
        v = cells / total / 1e9
        sample = "first %d alignments of the step batch per step, %d threads" % (n, threads)
        return ({"impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})

    import torch
    import ctypes as C
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (bbmap_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist          # initialised by the caller with the gloo backend: the data path has no collective
    msa = MultiStateAligner11tsCUDA(device=local, bandwidth=args.bandwidth, bandwidthRatio=args.ratio)
    if args.no_narrow:
        msa.set_option("narrow", 0)
    if args.no_strip:
        msa.set_option("strip", 0)
    if args.narrow_slack >= 0:
        msa.set_option("narrow", args.narrow_slack)
    if args.strip_buckets >= 0:
        msa.set_option("strip", args.strip_buckets)
    if os.environ.get("BBM_STRIP_STATS"):
        msa.set_option("strip_debug", 4)
    if args.strip_budget_mb:
        msa.set_option("strip_budget_mb", args.strip_budget_mb)
    dev = torch.device("cuda", local)
    # resident inputs (torch owns the device memory; the C ABI takes raw pointers)
    d_genome = torch.from_numpy(np.concatenate([genome, np.full(256, ord("N"), np.uint8)])).to(dev)
    d_reads = torch.from_numpy(reads).to(dev)
    d_tasks = torch.from_numpy(tasks.view(np.uint8)).to(dev)
    d_moff = torch.from_numpy(moff).to(dev)
    d_outs = torch.zeros(len(tasks) * wl.OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_match = torch.zeros(int(moff[-1]) + 16, dtype=torch.uint8, device=dev)
    max_rows = int(tasks["read_len"].max()); max_cols = int((tasks["ref_end"] - tasks["ref_start"] + 1).max())
    stream = torch.cuda.current_stream()

    def step_dev():
        return msa.align_batch_dev(d_reads.data_ptr(), d_genome.data_ptr(), d_tasks.data_ptr(), d_outs.data_ptr(), len(tasks),
                                   d_match.data_ptr(), d_moff.data_ptr(), max_rows, max_cols, C.c_void_p(stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # integer / DPX pipe peaks of this very GPU (the DP kernels' roofline denominators; not in MEASURED_PEAKS.json)
    kinds = ["iadd3", "lop3", "vimnmx3_dpx", "viaddmnmx_dpx", "imad", "half_imad_half_lop3", "cmp_select"]
    int_peaks = {k: msa.int_peak(i) for i, k in enumerate(kinds)} if rank == 0 else {}
    for _ in range(max(args.warmup, 3)):
        step_dev()
    outs = np.frombuffer(d_outs.cpu().numpy().tobytes(), dtype=wl.OUT_DTYPE)
    cells_per_step = int(outs["iterations"].sum())
    # one untimed step with the kernels' own work counters on: cells actually evaluated (strip kernel: rows of 8 cells; narrow kernel: 16
    # diagonals per row of every alignment that tried it), for the integer-issue roofline below
    msa.set_option("strip_debug", 4)
    u0, n0 = msa.stat("strip_units"), msa.stat("narrow_tried")
    step_dev()
    tried = msa.stat("narrow_tried") - n0
    evaluated_cells = 8 * (msa.stat("strip_units") - u0) + 16 * int(tasks["read_len"].mean()) * tried
    strip_lane_util = (msa.stat("strip_units") - u0) / max(1, msa.stat("strip_lane_iters"))
    msa.set_option("strip_debug", 0)
    assert (outs["status"] == 0).all(), "bench: some alignments returned an error status"
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    l0 = msa.launches
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kernel_ms = 0.0
    for _ in range(args.steps):
        kernel_ms += step_dev()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = msa.launches - l0
    # e2e: host buffers through the plug-in call, copies inside the timed region
    d_ref_ptr = C.c_void_p(d_genome.data_ptr())
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()           # pinned host buffers, as a production host would hold
    reads_p = pin(reads); tasks_p = pin(tasks.view(np.uint8)).view(wl.TASK_DTYPE); moff_p = pin(moff)
    outs_p = pin(np.zeros(len(tasks) * wl.OUT_DTYPE.itemsize, np.uint8)).view(wl.OUT_DTYPE)
    mbuf_p = pin(np.zeros(int(moff[-1]), np.int8))
    # Several batches in flight (default 3), the way the reference keeps one MSA per mapping thread (AbstractMapThread.java:133-136): one
    # host thread per batch, each with its own context, staging buffers and pinned result buffers, so the copies of one batch overlap
    # the kernels of the others.  Every step still pays its own H2D of tasks+reads and D2H of results+match strings inside the timed region.
    msa_b = MultiStateAligner11tsCUDA(device=local, bandwidth=args.bandwidth, bandwidthRatio=args.ratio)
    outs_q = pin(np.zeros(len(tasks) * wl.OUT_DTYPE.itemsize, np.uint8)).view(wl.OUT_DTYPE)
    mbuf_q = pin(np.zeros(int(moff[-1]), np.int8))
    lanes = [(msa, outs_p, mbuf_p), (msa_b, outs_q, mbuf_q)]
    nfl = int(os.environ.get("BBM_E2E_IN_FLIGHT", "3"))
    extra = []
    for _k in range(nfl - 2):
        extra.append(MultiStateAligner11tsCUDA(device=local, bandwidth=args.bandwidth, bandwidthRatio=args.ratio))
        lanes.append((extra[-1], pin(np.zeros(len(tasks) * wl.OUT_DTYPE.itemsize, np.uint8)).view(wl.OUT_DTYPE), pin(np.zeros(int(moff[-1]), np.int8))))
    for m_, o_, b_ in lanes:
        m_.align_batch(reads_p, d_ref_ptr, tasks_p, match_off=moff_p, outs=o_, mbuf=b_)
    barrier()
    e2e_steps = nfl * max(1, min(args.steps, 6) // nfl)

    def e2e_worker(k):
        m_, o_, b_ = lanes[k]
        for _ in range(e2e_steps // nfl):
            m_.align_batch(reads_p, d_ref_ptr, tasks_p, match_off=moff_p, outs=o_, mbuf=b_, account=False)

    workers = [threading.Thread(target=e2e_worker, args=(k,)) for k in range(nfl)]
    t0 = time.perf_counter()
    for w_ in workers:
        w_.start()
    for w_ in workers:
        w_.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h_outs = outs_p
    assert outs_q.tobytes() == outs.tobytes(), "second in-flight batch disagrees with the resident path"
    msa_b.close()
    for m_ in extra:
        m_.close()
    clocks = sampler.finish()
    assert h_outs.tobytes() == outs.tobytes(), "host-buffer path and resident path disagree"

    from bbmap_b200 import shard
    ms_all, e2e_ms_step, kernel_ms_all = shard.max_over_ranks([ms, e2e_s * 1e3 / e2e_steps, kernel_ms])
    total_cells_step, = shard.sum_over_ranks([float(cells_per_step)])
    if rank != 0:
        msa.close()
        return None
    value = total_cells_step * args.steps / (ms_all / 1e3) / 1e9
    e2e_value = total_cells_step / (e2e_ms_step / 1e3) / 1e9
    pk, pk_kind = peaks()
    # algorithmic HBM bytes per step: task + read + window + result + match string
    cols = (tasks["ref_end"] - tasks["ref_start"] + 1).astype(np.int64)
    alg_bytes = int(len(tasks) * (40 + 80) + tasks["read_len"].sum() + cols.sum() + np.maximum(outs["match_len"], 0).sum())
    step_s = ms_all / 1e3 / args.steps
    hbm_ach = alg_bytes / step_s / 1e9
    h2d = int(tasks.nbytes + reads.nbytes + moff.nbytes)
    d2h = int(outs.nbytes + moff[-1])
    line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": config,
            "alignments_per_s": len(tasks) * world / step_s,
            "computed_cells_gcups": float((tasks["read_len"].astype(np.int64) * cols).sum()) * world / step_s / 1e9,
            "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_step, "batches_in_flight": nfl},
            "gpu_launches": int(launches),
            "kernel_mix": {"tasks": msa.stat("tasks_total"), "narrow_tried": msa.stat("narrow_tried"),
                           "narrow_handed_over": msa.stat("narrow_handed_over"), "strip_tasks": msa.stat("strip_tasks"), "strip_units": msa.stat("strip_units"), "strip_lane_iters": msa.stat("strip_lane_iters"), "band_misses": msa.stat("band_misses")},
            "clocks": clocks,
            "int_peaks_glops": int_peaks,
            "roofline_int": {"bound": "integer ALU issue (compare/select/min-max/logic; 64 lanes/clk/SM)",
                             "evaluated_cells_per_s": evaluated_cells * world / step_s,
                             "floor_lane_ops_per_cell": 45, "achieved": evaluated_cells * world / step_s * 45 / 1e9,
                             "peak": 2 * int_peaks.get("cmp_select", 0.0) * world, "unit": "G lane-ops/s",
                             "frac": (evaluated_cells / step_s * 45 / 1e9) / max(1e-9, 2 * int_peaks.get("cmp_select", 0.0)),
                             "strip_lane_utilisation": strip_lane_util,
                             "note": "achieved = cells the kernels evaluate x the 45-op floor of the 3-state recurrence (SURVEY 8d); peak = measured "
                                     "compare+select issue rate of this GPU (bbm_int_peak); the kernels spend ~150 instructions per cell today"},
            "roofline": {"bound": "int-issue", "achieved": total_cells_step / step_s * 45 / 1e9, "peak": 2 * int_peaks.get("cmp_select", 0.0) * world,
                         "unit": "G lane-ops/s", "frac": (total_cells_step / world / step_s * 45 / 1e9) / max(1e-9, 2 * int_peaks.get("cmp_select", 0.0)),
                         "traffic": None, "kernel": "msa_strip_fill_kernel (+ narrow / prep / finish)",
                         "note": "achieved = REFERENCE cell visits/s x the 45-lane-op floor of the 3-state recurrence (SURVEY 8d); peak = 2 x the compare+select "
                                 "issue rate bbm_int_peak measured on this GPU in this run; roofline_int counts the cells the kernels actually evaluate",
                         "hbm": {"achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"], "peak_kind": pk_kind}}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_run(reads, genome, tasks, moff, args.bandwidth, args.ratio)
    try:
        msa.close()
    except Exception:
        pass
    return line


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist
        dist.init_process_group("gloo")
    line = measure(args, rank, world, local, reference=(args.impl == "reference"))
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
