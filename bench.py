#!/usr/bin/env python
"""bench.py — mapped reads/s of the whole seed-and-extend chain on BASELINE.json configs[1]-shaped input, one B200 per rank.

One "step" = one call of the batched mapper over one batch of synthetic 2x150 bp pairs — by default the 1 M pairs BASELINE configs[1] names, as ONE batch (E. coli-sized random reference, ~1 % substitutions,
1-3 bp indels, Q30): Read.validate -> KeyRing -> BBIndex.find -> pairSiteScoresInitial -> scoreNoIndels -> findTipDeletions -> scoreSlow -> rescue
-> pairSiteScoresFinal -> genMatchString / realign_new -> SamLine fields -> SAM text, i.e. BBMapThread.processReadPair for every pair.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--genome G] [--scaffolds S] [--impl reference]

`value`   reads/s with the reads resident in HBM (bbm_map_batch_dev), SAM text formatted on the device;
`e2e`     reads/s through the reference-facing call bbm_map_batch_host: FASTQ-shaped pinned host buffers in, SAM text + records out, every
          copy inside the timed region;
`roofline` the dominant kernels of the step — the MultiStateAligner11ts fills — against the integer-issue roof measured in the same run;
`msa`     the MultiStateAligner11ts microbenchmark (configs[2], GCUPS; bench/msa_bench.py), `msa_band_sweep` its band sweep (bw 12 / 40, ratio 0.18),
          `banded` BandedAligner at G6 scale, `spliced` RNA-seq-style pairs with 0.3-12 kbp introns through the same mapper call (configs[4] shape within the
          default maxindel, intronlen=10), `long_reads` 1-kbp single-ended reads cut at 500 by bbm_break_reads (maxlen=500) and mapped piece by piece;
`cpu_baseline` / `--impl reference`: the same chain through the sequential CPU restatement (oracle/), one process per host core.
Prints ONE JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bench"))

from bbmap_b200 import workloads as wl  # noqa: E402

METRIC = "mapped reads/sec (2x150bp pairs, SAM text out)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per step per GPU (configs[1] names 1 M pairs: one step maps all of them as one batch)")
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--scaffolds", type=int, default=1)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-msa", action="store_true", help="skip the MSA microbenchmark (configs[2]) sub-measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the band sweep / BandedAligner / single-ended sub-measurements")
    ap.add_argument("--msa-tasks", type=int, default=1_600_000)
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs of the CPU sample (0 = sized for ~20 s)")
    ap.add_argument("--in-flight", type=int, default=2, help="mapper contexts of the end-to-end arm (batches in flight)")
    ap.add_argument("--opt", action="append", default=[], help="bbm_set_option key=value applied to every mapper context (A/B while tuning)")
    return ap.parse_args()


def make_reference(genome_len, scaffolds):
    """G2: one random scaffold; G3: `scaffolds` random scaffolds of unequal size + 1 % of the bases in planted 300-bp repeat families of 100 copies."""
    if scaffolds <= 1:
        return [wl.random_genome(genome_len, seed=1)]
    rng = np.random.Generator(np.random.PCG64(3))
    w = rng.uniform(0.4, 2.0, size=scaffolds); sizes = np.maximum(1000, (w / w.sum() * genome_len).astype(np.int64))
    scafs = [wl.ACGT[rng.integers(0, 4, size=int(z), dtype=np.uint8)] for z in sizes]
    fam = max(1, int(genome_len * 0.01 / 300 / 100))
    for _ in range(fam):
        unit = wl.ACGT[rng.integers(0, 4, size=300, dtype=np.uint8)]
        for _c in range(100):
            sc = scafs[int(rng.integers(0, scaffolds))]
            q = int(rng.integers(0, len(sc) - 300)); sc[q:q + 300] = unit
    return scafs


def read_names(n):
    """FASTQ-style names: pair index and /1 /2."""
    names = [b"p%d/%d" % (i // 2, (i & 1) + 1) for i in range(n)]
    off = np.zeros(n + 1, np.int64); np.cumsum([len(x) for x in names], out=off[1:])
    return np.frombuffer(b"".join(names), np.int8).copy(), off, names


# ---------------------------------------------------------------------------------------------------------------------------------
# CPU arm: the sequential restatement (oracle/chain.py), one worker process per host core, each mapping its own slice of the sample
_W = {}


def _cpu_worker(args):
    lo, hi, paired = args
    from oracle import chain, oracle as orc
    o = orc.get()
    R = _W["R"]; L = 150
    a, b = 2 * lo * L, 2 * hi * L
    off = np.arange(2 * (hi - lo) + 1, dtype=np.int64) * L
    t0 = time.perf_counter()
    fn = chain.map_pairs if paired else chain.map_single
    res = fn(o, _W["idx"], _W["cb"], _W["co"], _W["table"], R["bases"][a:b], R["qual"][a:b], off)
    lines = chain.sam_lines(res, off, names=_W["names"][2 * lo:2 * hi], scaf_names=_W["snames"], paired=paired)
    return time.perf_counter() - t0, int((res["recs"]["flags"] & 1).sum()), sum(len(x) for x in lines)


def cpu_chain(cb, co, table, R, names, snames, npairs, threads, paired=True):
    """Maps the first `npairs` pairs with `threads` worker processes (fork: the index is built once and shared copy-on-write)."""
    import multiprocessing as mp
    from oracle import oracle as orc
    o = orc.get()
    if _W.get("cb_id") != id(cb):
        _W.update(idx=o.index_build(cb, co, 13, -1), cb_id=id(cb))          # built once, outside the timed region (the reference loads its index once per run)
    _W.update(R=R, cb=cb, co=co, table=table, names=names, snames=snames)
    threads = max(1, min(threads, npairs))
    bounds = [(npairs * k // threads, npairs * (k + 1) // threads, paired) for k in range(threads)]
    t0 = time.perf_counter()
    if threads == 1:
        outs = [_cpu_worker(bounds[0])]
    else:
        with mp.get_context("fork").Pool(threads) as pool:
            outs = pool.map(_cpu_worker, bounds)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "reads": 2 * npairs, "reads_per_s": 2 * npairs / dt, "mapped": sum(x[1] for x in outs), "sam_bytes": sum(x[2] for x in outs), "threads": threads}


def reference_arm(args, cb, co, table, R, names, snames):
    threads = os.cpu_count() or 1
    npairs = args.cpu_pairs or min(args.pairs, threads * 6000)
    times, reads = [], 0
    for it in range(args.warmup + args.steps):
        r = cpu_chain(cb, co, table, R, names, snames, npairs, threads)
        if it >= args.warmup:
            times.append(r["seconds"]); reads += r["reads"]
    v = reads / sum(times)
    sample = "the first %d pairs of the step batch per step, %d worker processes, sequential C restatement of the chain (oracle/) incl. SAM text" % (npairs, threads)
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_of(args), "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "no JVM in the image: the reference's Java host cannot run; this is its restatement in C (the fills are pinned to the reference's own C by tests/test_oracle_vs_reference.py)"}


def config_of(args):
    return {"workload": "configs[1]: %d bp random reference (%d scaffold(s)), %d pairs of 2x150 bp per step per GPU, insert 200-500, ~1%% substitutions, 1-3 bp indel in ~50%% of the reads, "
                        "Q30; BBMap default flags (k=13, ambig=best, rescue on), index and reference resident and replicated per GPU" % (args.genome, args.scaffolds, args.pairs),
            "pairs_per_step_per_gpu": args.pairs, "genome_bp": args.genome, "scaffolds": args.scaffolds,
            "l2": "inputs larger than L2 (%.0f MB of reads + %.0f MB of site lists per step)" % (args.pairs * 2 * 300 / 1e6, args.pairs * 2 * 16 * 80 / 1e6)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    scafs = make_reference(args.genome, args.scaffolds)
    snames_all = ["scaffold_%d" % (i + 1) for i in range(len(scafs))]
    from bbmap_b200.index import pack_chromosomes
    if args.impl == "reference":
        if rank != 0:
            return
        cb, co, table = pack_chromosomes(scafs)
        npairs = args.cpu_pairs or min(args.pairs, (os.cpu_count() or 1) * 6000)
        R = wl.make_mapping_reads(cb, co, table, npairs, seed=2)
        _, _, names = read_names(2 * npairs)
        order = sorted(range(len(table)), key=lambda i: table[i])
        print(json.dumps(reference_arm(args, cb, co, table, R, names, [snames_all[i].encode() for i in order])))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (bbmap_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")            # barrier + max-over-ranks only: the data path has no collective, so no NCCL
    import ctypes as C
    from bbmap_b200 import lib as _lib, shard
    from bbmap_b200.mapper import BBMapCUDA, MAP_REC_DTYPE, MAP_STATS_DTYPE, mapper_cfg
    from bbmap_b200.sam import SAM_OUT_DTYPE
    from msa_bench import ClockSampler, peaks
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    t_build = time.perf_counter()
    m = BBMapCUDA(scafs, names=snames_all, device=local)
    t_build = time.perf_counter() - t_build
    del scafs
    cb, co, table = m.cb, m.co, m.table
    R = wl.make_mapping_reads(cb, co, table, args.pairs, seed=2 + rank)
    n = 2 * args.pairs
    nbuf, noff, names = read_names(n)
    cfg = mapper_cfg(paired=True, sam_text=True)
    L = m.L; h = m.h
    L.bbm_set_option(h, b"msa_count", 0)
    opts = [(kv.split("=")[0].encode(), int(kv.split("=")[1])) for kv in args.opt]
    for k_, v_ in opts:
        _lib.check(L.bbm_set_option(h, k_, v_), "bbm_set_option")

    # ---- resident arm ----
    pad = lambda a, extra=64: torch.from_numpy(np.concatenate([a, np.zeros(extra, a.dtype)])).to(dev)
    d_bases = pad(R["bases"]); d_qual = pad(R["qual"]); d_off = torch.from_numpy(R["off"]).to(dev)
    d_recs = torch.zeros(n * MAP_REC_DTYPE.itemsize, dtype=torch.uint8, device=dev); d_sam = torch.zeros(n * SAM_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    stats = np.zeros(1, MAP_STATS_DTYPE)
    stream = torch.cuda.current_stream()
    p = lambda t: C.c_void_p(t.data_ptr())

    def step_dev():
        _lib.check(L.bbm_map_batch_dev(h, p(d_bases), p(d_qual), p(d_off), n, 150, cfg.ctypes.data_as(C.c_void_p), p(d_recs), p(d_sam), None, 0,
                                       C.c_void_p(stream.cuda_stream), stats.ctypes.data_as(C.c_void_p)), "bbm_map_batch_dev")
        return stats[0].copy()

    for _ in range(max(args.warmup, 3)):
        st0 = step_dev()
    # one untimed step with the aligner's cell counter on (roofline numerator) and the integer peaks of this very GPU (denominator)
    L.bbm_set_option(h, b"msa_count", 1)
    c0 = L.bbm_get_stat(h, b"msa_cells"); u0 = L.bbm_get_stat(h, b"msa_us")
    step_dev()
    cells_step = L.bbm_get_stat(h, b"msa_cells") - c0
    L.bbm_set_option(h, b"msa_count", 0)
    kinds = ["iadd3", "lop3", "vimnmx3_dpx", "viaddmnmx_dpx", "imad", "half_imad_half_lop3", "cmp_select"]
    int_peaks = {}
    if rank == 0:
        for i, k in enumerate(kinds):
            g = C.c_double(0); _lib.check(L.bbm_int_peak(h, i, C.byref(g)), "bbm_int_peak"); int_peaks[k] = g.value
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.3)
    barrier()
    l0 = L.bbm_launch_count(h); u0 = L.bbm_get_stat(h, b"msa_us")
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    stage = np.zeros(7)
    for _ in range(args.steps):
        s_ = step_dev()
        stage += [s_["ms_seed_search"], s_["ms_lists"], s_["ms_slow"], s_["ms_rescue"], s_["ms_genmatch"], s_["ms_sam"], s_["ms_total"]]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = L.bbm_launch_count(h) - l0
    msa_ms_step = (L.bbm_get_stat(h, b"msa_us") - u0) / 1e3 / args.steps
    recs = np.frombuffer(d_recs.cpu().numpy().tobytes(), MAP_REC_DTYPE)
    tr = R["truth"]
    mapped = (recs["flags"] & 1) != 0
    correct = mapped & (recs["chrom"] == tr[:, 0]) & (recs["strand"] == tr[:, 1]) & ((np.abs(recs["start"] - tr[:, 2]) <= 8) | (np.abs(recs["stop"] - tr[:, 3]) <= 8))
    exact = mapped & (recs["start"] == tr[:, 2]) & (recs["stop"] == tr[:, 3])

    # ---- end to end: host buffers through the reference-facing call, `in_flight` batches at a time on their own contexts / host threads ----
    import threading
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    hb = pin(R["bases"].view(np.int8)); hq = pin(R["qual"].view(np.int8)); ho = pin(R["off"]); hn = pin(nbuf); hno = pin(noff)
    sam_cap = int(st0["sam_bytes"]) + int(nbuf.nbytes) + (1 << 20)          # the resident arm had no read names (QNAME "*")
    nfl = max(1, args.in_flight)
    lanes = []
    for k in range(nfl):
        mk = m if k == 0 else m.clone()
        for k_, v_ in opts:
            L.bbm_set_option(mk.h, k_, v_)
        lanes.append({"m": mk, "recs": pin(np.zeros(n * MAP_REC_DTYPE.itemsize, np.int8)), "sam": pin(np.zeros(n * SAM_OUT_DTYPE.itemsize, np.int8)),
                      "text": pin(np.zeros(sam_cap, np.int8)), "toff": pin(np.zeros(n + 1, np.int64)), "stats": np.zeros(1, MAP_STATS_DTYPE)})

    def step_host(lane):
        _lib.check(L.bbm_map_batch_host(lane["m"].h, hb.ctypes.data_as(C.c_void_p), hq.ctypes.data_as(C.c_void_p), ho.ctypes.data_as(C.c_void_p), n, hn.ctypes.data_as(C.c_void_p),
                                        hno.ctypes.data_as(C.c_void_p), cfg.ctypes.data_as(C.c_void_p), lane["recs"].ctypes.data_as(C.c_void_p), lane["sam"].ctypes.data_as(C.c_void_p),
                                        None, 0, lane["text"].ctypes.data_as(C.c_void_p), sam_cap, lane["toff"].ctypes.data_as(C.c_void_p), lane["stats"].ctypes.data_as(C.c_void_p)),
                   "bbm_map_batch_host")

    for lane in lanes:
        step_host(lane)
    barrier()
    per_lane = max(1, args.steps // nfl)
    e2e_steps = per_lane * nfl

    def worker(lane):
        for _ in range(per_lane):
            step_host(lane)

    ths = [threading.Thread(target=worker, args=(lane,)) for lane in lanes]
    t0 = time.perf_counter()
    for t_ in ths:
        t_.start()
    for t_ in ths:
        t_.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.finish()
    sam_bytes = int(lanes[0]["stats"][0]["sam_bytes"])
    assert lanes[0]["recs"].tobytes() == recs.tobytes(), "host-buffer path and resident path disagree"
    for lane in lanes[1:]:
        assert lane["recs"].tobytes() == recs.tobytes() and lane["text"][:sam_bytes].tobytes() == lanes[0]["text"][:sam_bytes].tobytes(), "in-flight contexts disagree"
        lane["m"].close()
    first_lines = lanes[0]["text"][: int(lanes[0]["toff"][2])].tobytes().decode()

    ms_all, e2e_ms_step, msa_ms_all = shard.max_over_ranks([ms, e2e_s * 1e3 / e2e_steps, msa_ms_step])
    reads_all, mapped_all, cells_all = shard.sum_over_ranks([float(n), float(mapped.sum()), float(cells_step)])
    if rank != 0:
        m.close()
        return
    step_s = ms_all / 1e3 / args.steps
    value = reads_all / step_s
    e2e_value = reads_all / (e2e_ms_step / 1e3)
    pk, pk_kind = peaks()
    h2d = int(R["bases"].nbytes + R["qual"].nbytes + R["off"].nbytes + nbuf.nbytes + noff.nbytes)
    d2h = int(n * (MAP_REC_DTYPE.itemsize + SAM_OUT_DTYPE.itemsize) + sam_bytes + (n + 1) * 8)
    cmp_peak = 2 * int_peaks.get("cmp_select", 0.0)
    dp_s = msa_ms_all / 1e3
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config_of(args),
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_step, "batches_in_flight": nfl,
                    "sam_bytes_per_step": sam_bytes},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "int-issue", "kernel": "MultiStateAligner11ts fills of the step (scoreSlow + slowRescue + realign_new: classify, narrow, strip prep/fill/finish, tiled, generic)",
                         "achieved": cells_all / max(dp_s, 1e-9) * 45 / 1e9 / world, "peak": cmp_peak, "unit": "G lane-ops/s",
                         "frac": (cells_all / world / max(dp_s, 1e-9) * 45 / 1e9) / max(cmp_peak, 1e-9), "traffic": None,
                         "dp_ms_per_step": msa_ms_all, "dp_share_of_step": msa_ms_all / (ms_all / args.steps), "reference_cells_per_step_per_gpu": cells_all / world,
                         "dp_gcups": cells_all / world / max(dp_s, 1e-9) / 1e9,
                         "note": "achieved = reference cell visits of every alignment of the step (the reference's iterations counter) x the 45-lane-op floor of the 3-state "
                                 "recurrence / the device time of the aligner batches (CUDA events); peak = 2 x the compare+select issue rate measured by bbm_int_peak in this run",
                         "hbm_peak_gbs": pk["hbm_gbs"], "hbm_peak_kind": pk_kind},
            "int_peaks_glops": int_peaks,
            "stage_ms": dict(zip(["seed_search", "lists", "scoreSlow", "rescue_pairing", "genMatchString_finish", "sam", "total"], (stage / args.steps).tolist())),
            "mapping": {"mapped": float(mapped_all / reads_all), "paired": float(((recs["flags"] & 8) != 0).mean()), "rescued": float(((recs["flags"] & 16) != 0).mean()),
                        "ambiguous": float(((recs["flags"] & 4) != 0).mean()), "top_site_is_origin": float(correct.mean()), "exact_start_and_stop": float(exact.mean()),
                        "status_reads": int(st0["status_reads"]), "site_overflow_reads": int(st0["site_overflow_reads"]), "max_sites_used": int(st0["max_sites_used"]),
                        "slow_alignments": int(st0["slow_alignments"]), "rescue_scans": int(st0["rescue_scans"]), "rescue_fills": int(st0["rescue_fills"]),
                        "realign_fills": int(st0["realign_fills"]), "genmatch_rounds": int(st0["genmatch_rounds"]), "mated_pairs": int(st0["mated_pairs"]),
                        "mean_inner_length": float(st0["inner_length_sum"]) / max(1, int(st0["mated_pairs"])), "first_sam_lines": first_lines},
            "index_build_s": t_build}
    if not args.no_extras and world == 1:
        # configs[4] shape within the default maxindel (16000): RNA-seq-style pairs, half of them with one read spliced over a 0.3-12 kbp intron
        # (gapped candidate sites from BBIndex, makeGref fills, match strings with the intron as a D run); through the host-buffer call
        sp_pairs = min(args.pairs, 50_000)
        RS = wl.make_spliced_reads(cb, co, table, sp_pairs, seed=7)
        ns_ = 2 * sp_pairs
        cfg_s = mapper_cfg(paired=True, sam_text=False, match_slot=12288 + 384)
        cfg_s["sam"]["intron_limit"] = 10                 # `intronlen=10`: deletions longer than 10 are written as N in the CIGAR (SamLine.INTRON_LIMIT)
        ms_s = 12288 + 384
        for _ in range(2):
            dev_s = m.map_batch(RS["bases"], RS["qual"], RS["off"], cfg=cfg_s, match_stride=0)
        t0 = time.perf_counter(); reps_s = 3
        for _ in range(reps_s):
            dev_s = m.map_batch(RS["bases"], RS["qual"], RS["off"], cfg=cfg_s, match_stride=0)
        dt_s = (time.perf_counter() - t0) / reps_s
        rs_, ts_ = dev_s["recs"], RS["truth"]
        mp_ = (rs_["flags"] & 1) != 0
        ok_ = mp_ & (rs_["chrom"] == ts_[:, 0]) & (rs_["strand"] == ts_[:, 1]) & (np.abs(rs_["start"] - ts_[:, 2]) <= 8) & (np.abs(rs_["stop"] - ts_[:, 3]) <= 8)
        line["spliced"] = {"workload": "%d pairs of 2x150 bp on the same reference, one read of every second pair spliced over an intron of 300-12000 bp; default flags (maxindel 16000) + intronlen=10, "
                                       "bbm_map_cfg.match_slot %d" % (sp_pairs, ms_s),
                           "e2e_reads_per_s": ns_ / dt_s, "ms_per_batch": dt_s * 1e3, "mapped": float(mp_.mean()), "start_and_stop_within_8": float(ok_.mean()),
                           "spliced_reads": int(RS["spliced"].sum()), "spliced_start_and_stop_within_8": float(ok_[RS["spliced"]].mean()),
                           "status_reads": int((rs_["status"] != 0).sum()), "realign_fills": int(dev_s["stats"]["realign_fills"]), "slow_alignments": int(dev_s["stats"]["slow_alignments"])}
        # configs[4], second half: 1 kbp single-ended reads, cut into 500-base pieces by the host as `maxlen=500` does (ReformatReads.breakReads in
        # AbstractMapThread.run), every piece mapped as a read of its own through the same host-buffer call
        from bbmap_b200.reads import break_reads
        nlong = min(args.pairs, 20_000)
        RL = wl.make_long_reads(cb, co, table, nlong, L=1000, seed=8)
        cfg_l = mapper_cfg(paired=False, sam_text=False)
        t0 = time.perf_counter()
        P = break_reads(RL["bases"], RL["qual"], RL["off"], RL["names"], RL["name_off"], 500, 0)
        t_break = time.perf_counter() - t0
        for _ in range(2):
            dev_l = m.map_batch(P["bases"], P["quality"], P["read_off"], cfg=cfg_l, match_stride=0)
        t0 = time.perf_counter(); reps_l = 3
        for _ in range(reps_l):
            dev_l = m.map_batch(P["bases"], P["quality"], P["read_off"], cfg=cfg_l, match_stride=0)
        dt_l = (time.perf_counter() - t0) / reps_l
        rl_, tl_ = dev_l["recs"], RL["truth"][P["src"]]
        plen = np.diff(P["read_off"])
        mp_l = (rl_["flags"] & 1) != 0
        # a piece that starts at read offset o lies o bases into the footprint on the plus strand, o bases before its end on the minus strand (+- the indel)
        exp_start = np.where(tl_[:, 1] == 0, tl_[:, 2] + P["piece_start"], tl_[:, 3] - P["piece_start"] - plen + 1)
        ok_l = mp_l & (rl_["chrom"] == tl_[:, 0]) & (rl_["strand"] == tl_[:, 1]) & (np.abs(rl_["start"] - exp_start) <= 4) & (np.abs((rl_["stop"] - rl_["start"] + 1) - plen) <= 4)
        line["long_reads"] = {"workload": "%d single-ended reads of 1000 bp on the same reference (~1%% substitutions, a 1-3 bp indel in half of them), cut into %d pieces of <= 500 bp "
                                          "by bbm_break_reads (maxlen=500) and mapped through bbm_map_batch_host, default flags" % (nlong, len(plen)),
                              "e2e_long_reads_per_s": nlong / (dt_l + t_break), "e2e_pieces_per_s": len(plen) / (dt_l + t_break), "ms_per_batch": dt_l * 1e3, "break_ms": t_break * 1e3,
                              "mapped": float(mp_l.mean()), "piece_at_its_origin_within_4": float(ok_l.mean()), "status_reads": int((rl_["status"] != 0).sum()),
                              "slow_alignments": int(dev_l["stats"]["slow_alignments"]), "realign_fills": int(dev_l["stats"]["realign_fills"])}
    m.close()
    if not args.no_cpu_baseline and world == 1:          # the CPU figure is reported at N=1 only
        threads = os.cpu_count() or 1
        npairs = args.cpu_pairs or min(args.pairs, threads * 6000)
        order = sorted(range(len(table)), key=lambda i: table[i])
        cr = cpu_chain(cb, co, table, R, names, [snames_all[i].encode() for i in order], npairs, threads)
        line["cpu_baseline"] = {"value": cr["reads_per_s"], "unit": "reads/s", "cores": cr["threads"], "kind": "port",
                                "sample": "first %d pairs of the step batch, %d worker processes, sequential C restatement of the chain incl. SAM text, %.1f s" % (npairs, cr["threads"], cr["seconds"]),
                                "mapped": cr["mapped"] / cr["reads"]}
    if not args.no_msa:
        import msa_bench
        a2 = argparse.Namespace(gpus=args.gpus, steps=min(args.steps, 5), warmup=3, tasks=args.msa_tasks, impl="cuda", bandwidth=0, ratio=0.0, no_cpu_baseline=False,
                                lengths="100,150,250", no_narrow=False, no_strip=False, strip_budget_mb=0, narrow_slack=-1, strip_buckets=-1, no_stages=True, stage_pairs=0)
        if world == 1:
            line["msa"] = msa_bench.measure(a2, 0, 1, local)
            if not args.no_extras:
                sweep = []
                for bw, ratio in ((12, 0.0), (40, 0.0), (0, 0.18)):
                    a3 = argparse.Namespace(**vars(a2)); a3.bandwidth = bw; a3.ratio = ratio; a3.no_cpu_baseline = True; a3.tasks = min(args.msa_tasks, 400_000); a3.steps = 3
                    r3 = msa_bench.measure(a3, 0, 1, local)
                    sweep.append({"bandwidth": bw, "bandwidthRatio": ratio, "gcups": r3["value"], "e2e_gcups": r3["e2e"]["value"], "ms_per_step": r3["ms_per_step"],
                                  "alignments_per_s": r3["alignments_per_s"], "band_misses": r3["kernel_mix"]["band_misses"]})
                line["msa_band_sweep"] = sweep
    if not args.no_extras and world == 1:
        import banded_bench
        line["banded"] = banded_bench.run(device=local)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
