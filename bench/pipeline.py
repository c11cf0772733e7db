"""The device stages of the unpaired mapping loop chained with everything resident in HBM (no host round trip between stages):

    Read.validate -> KeyRing seeds -> BBIndex.find -> SiteScore lists -> removeOutOfBounds -> trimList -> scoreNoIndels(Read) -> scoreSlow (rounds over the
    MultiStateAligner11ts kernels, padding retry included; findTipDeletions before it) -> mergeDuplicateSites / clearzone / removeLowQualitySitesUnpaired

on BASELINE configs[1]-shaped input (E. coli-sized random reference, 2x150 bp reads mapped as single reads, ~1 % substitutions, 1-3 bp
indels, Q30).  What the reference's processRead (current/align2/BBMapThread.java:389-733) does in addition and is NOT in this chain yet:
genMatchString/realign_new, applyClearzone3, the tip-score penalty, pairing and rescue, SAM
text.  The reads/s printed here is therefore the throughput of the built device stages up to the final site decision (strand, POS, score,
ambiguity), not yet a whole-mapper number.

    python bench/pipeline.py [--pairs 200000] [--genome 4600000] [--reps 3]

Prints one JSON object; bench.py embeds it as `pipeline`."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bbmap_b200 import lib as _lib, sitelist as sl, workloads as wl  # noqa: E402
from bbmap_b200.index import BBIndexCUDA, pack_chromosomes  # noqa: E402
from bbmap_b200.keyring import default_cfg  # noqa: E402
from bbmap_b200.search import HEAD_DTYPE, SITE_DTYPE  # noqa: E402

MAXK, MAX_SITES, CAP = 32, 16, 16


def cpu_side(R, cb, co, table, n_cpu=20000):
    """The same chain through the CPU oracle (C restatements, one host thread) on the first n_cpu reads — a reported baseline."""
    from oracle import oracle as orc
    o = orc.get()
    m = min(n_cpu, len(R["off"]) - 1)
    off = R["off"][:m + 1]; nb = int(off[-1])
    idx = o.index_build(cb, co, 13, -1)                       # not timed: the index is built once per genome
    t0 = time.perf_counter()
    bases, qual, basesM, _ = o.ingest_batch(R["bases"][:nb], R["qual"][:nb], off, 0)
    seeds = o.seed_batch(bases, qual, off, default_cfg(), MAXK)
    res = o.search_batch(idx, cb, co, bases, seeds["baseScores"], off, seeds, quit_after_two_perfects=False)
    ns = np.minimum(res["nsites"], CAP).astype(np.int32)
    lists = np.zeros((m, CAP), sl.SS_DTYPE); S = res["sites"][:, :CAP]
    for f in ("chrom", "start", "stop", "hits", "score", "strand", "perfect", "semiperfect", "ngaps", "gaps"):
        lists[f] = S[f]
    lists["quick_score"] = S["score"]
    pcfg = sl.policy_cfg()
    from bbmap_b200.sam import scaffold_table
    lists, ns, _ = o.sitelist_bounds(lists, ns, off, (np.diff(np.asarray(co, np.int64)) - 1).astype(np.int32), scaffold_table(table, len(co) - 1))
    lists, ns, _ = o.sitelist(sl.SL_TRIM, lists, ns, off, pcfg)
    lists, ns, out = o.sitelist(sl.SL_NOINDEL, lists, ns, off, pcfg, bases, basesM, cb, co)
    runm = (out["near_perfect"] < 1).astype(np.int32)
    from bbmap_b200.rescue import tipdel_cfg
    lists, _ = o.sitelist_tipdel(lists, ns * runm, off, bases, basesM, qual, cb, co, tipdel_cfg())
    lists, _, na = o.score_slow(lists, ns, off, bases, basesM, cb, co, runm, sl.slow_cfg())
    lists, ns, out = o.sitelist(sl.SL_FINAL, lists, ns, off, pcfg)
    lists, ns, out = o.sitelist_clearzone3(lists, ns, off, out, pcfg)
    from bbmap_b200 import sam as _sam
    srec, _, _ = o.sam_batch(_sam.tasks_from_lists(lists, ns, off, out), np.zeros(1, np.int8), scaffold_table(table, len(co) - 1), _sam.default_cfg())
    dt = time.perf_counter() - t0
    return {"reads": m, "cores": 1, "kind": "port", "seconds": dt, "reads_per_s": m / dt, "slow_alignments": int(na),
            "mapped": float(((out["flags"] & sl.F_MAPPED) != 0).mean()), "lists": lists, "nss": ns, "flags": out["flags"].copy(), "sam": srec}


def make_reference(genome_len, scaffolds):
    """One random scaffold (G2), or G3 of SURVEY 8d: `scaffolds` random scaffolds of unequal size + 1 % of the bases in planted 300-bp repeat
    families of 100 copies (same generator as bench/stages.py)."""
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


def run(pairs=200_000, genome_len=4_600_000, reps=3, device=0, cpu=True, seed=2, scaffolds=1):
    import torch
    L = _lib.load()
    dev = torch.device("cuda", device)
    torch.cuda.set_device(device)
    scafs = make_reference(genome_len, scaffolds)
    cb, co, table = pack_chromosomes(scafs)
    del scafs
    R = wl.make_mapping_reads(cb, co, table, pairs, seed=seed)
    n = 2 * pairs; nb = len(R["bases"])
    idx = BBIndexCUDA(cb, co, keylen=13, device=device)
    h = idx.h
    p = lambda t: C.c_void_p(t.data_ptr())
    pad = lambda a, extra=64: torch.from_numpy(np.concatenate([a, np.zeros(extra, a.dtype)])).to(dev)
    d_bases0 = pad(R["bases"]); d_qual0 = pad(R["qual"]); d_off = torch.from_numpy(R["off"]).to(dev)
    d_bases = torch.empty_like(d_bases0); d_qual = torch.empty_like(d_qual0)
    d_basesM = torch.zeros(nb + 64, dtype=torch.uint8, device=dev); d_flags = torch.zeros(n, dtype=torch.int32, device=dev)
    d_nkeys = torch.zeros(n, dtype=torch.int32, device=dev)
    d_offsets = torch.zeros(n * MAXK, dtype=torch.int32, device=dev); d_keys = torch.zeros_like(d_offsets); d_ks = torch.zeros_like(d_offsets)
    d_offM = torch.zeros_like(d_offsets); d_keysM = torch.zeros_like(d_offsets)
    d_bs = torch.zeros(nb + 64, dtype=torch.int8, device=dev)
    d_heads = torch.zeros(n * HEAD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_sites = torch.zeros(n * MAX_SITES * SITE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_lists = torch.zeros(n * CAP * sl.SS_DTYPE.itemsize, dtype=torch.uint8, device=dev); d_nss = torch.zeros(n, dtype=torch.int32, device=dev)
    d_out = torch.zeros(n * sl.READ_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_co = torch.from_numpy(np.ascontiguousarray(co, np.int64)).to(dev)
    d_chroms = C.c_void_p(idx.d_chroms.value)
    from bbmap_b200.sam import scaffold_table
    so_, sl_, _sn = scaffold_table(table, len(co) - 1)
    d_scaf_off = torch.from_numpy(np.ascontiguousarray(so_, np.int32)).to(dev); d_scaf_loc = torch.from_numpy(np.ascontiguousarray(sl_, np.int32)).to(dev)
    d_maxidx = torch.from_numpy((np.diff(np.asarray(co, np.int64)) - 1).astype(np.int32)).to(dev)       # ChromosomeArray.maxIndex of the packed arrays
    from bbmap_b200.rescue import tipdel_cfg
    scfg = default_cfg(); pcfg = sl.policy_cfg(); wcfg = sl.slow_cfg(); tcfg = tipdel_cfg()
    d_out2 = torch.zeros(n * sl.READ_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    from bbmap_b200 import sam as _sam
    samcfg = _sam.default_cfg()
    d_scaf_len = torch.from_numpy(np.ascontiguousarray(_sn, np.int32)).to(dev)
    d_stasks = torch.zeros(n * _sam.SAM_TASK_DTYPE.itemsize, dtype=torch.uint8, device=dev); d_souts = torch.zeros(n * _sam.SAM_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_nomatch = torch.zeros(16, dtype=torch.int8, device=dev); d_coff = (torch.arange(n + 1, dtype=torch.int64, device=dev) * 12).contiguous()   # 2*match_len+12 per record
    d_cigar = torch.zeros(12 * n + 16, dtype=torch.int8, device=dev)
    cp = lambda a: a.ctypes.data_as(C.c_void_p)
    ms = C.c_float(0); na = C.c_int64(0)
    t = {}

    def step(name, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        t[name] = t.get(name, []) + [1e3 * (time.perf_counter() - t0)]

    def sitelist(op):
        _lib.check(L.bbm_sitelist_batch_dev(h, op, p(d_lists), p(d_nss), n, CAP, p(d_off), p(d_bases), p(d_basesM), d_chroms, p(d_co), cp(pcfg), p(d_out), None, None),
                   "bbm_sitelist_batch_dev")

    run_flags = [None]; masked = [None]

    def chain():
        d_bases.copy_(d_bases0); d_qual.copy_(d_qual0)            # Read.validate works in place
        step("ingest", lambda: _lib.check(L.bbm_ingest_batch_dev(h, p(d_bases), p(d_qual), p(d_off), n, 150, 0, p(d_basesM), p(d_flags), None, C.byref(ms)), "ingest"))
        step("seed", lambda: _lib.check(L.bbm_seed_batch_dev(h, p(d_bases), p(d_qual), p(d_off), n, 150, cp(scfg), MAXK, p(d_nkeys), p(d_offsets), p(d_keys), p(d_ks),
                                                             p(d_bs), p(d_offM), p(d_keysM), None, C.byref(ms)), "seed"))
        step("search", lambda: _lib.check(L.bbm_search_batch_dev(h, p(d_bases), p(d_bs), p(d_off), n, p(d_nkeys), p(d_offsets), p(d_ks), MAXK, 0, p(d_heads), p(d_sites),
                                                                 MAX_SITES, 150, None, C.byref(ms)), "search"))
        step("lists", lambda: _lib.check(L.bbm_sitelist_from_search_dev(h, p(d_heads), p(d_sites), n, MAX_SITES, p(d_lists), p(d_nss), CAP, None), "from_search"))
        step("removeOutOfBounds", lambda: _lib.check(L.bbm_sitelist_bounds_dev(h, p(d_lists), p(d_nss), n, CAP, p(d_off), p(d_maxidx), p(d_scaf_off), p(d_scaf_loc), 300, 1,
                                                                              2522, p(d_out2), None), "bounds"))
        step("trimList", lambda: sitelist(sl.SL_TRIM))
        step("scoreNoIndels", lambda: sitelist(sl.SL_NOINDEL))

        step("run_mask", lambda: run_flags.__setitem__(0, (d_out.view(torch.int32).view(n, 4)[:, 0] < 1).to(torch.int32).contiguous()))   # processRead :455-465
        # findTipDeletions runs under the same condition; reads that skip it keep nss as is (the kernel sees 0 sites for them)
        masked[0] = (d_nss * run_flags[0]).contiguous()
        step("findTipDeletions", lambda: _lib.check(L.bbm_sitelist_tipdel_dev(h, p(d_lists), p(masked[0]), n, CAP, p(d_off), p(d_bases), p(d_basesM),
                                                                             p(d_qual), d_chroms, p(d_co), None, cp(tcfg), p(d_out2), None, None), "tipdel"))

        def slow():
            _lib.check(L.bbm_scoreslow_dev(h, p(d_lists), p(d_nss), n, CAP, p(d_off), p(d_bases), p(d_basesM), d_chroms, p(d_co), p(run_flags[0]), cp(wcfg), p(d_status),
                                           150, None, C.byref(na), None), "bbm_scoreslow_dev")
        step("scoreSlow", slow)
        step("final", lambda: sitelist(sl.SL_FINAL))
        # processRead :667-700 (in the reference the primary site's match string is generated in between; it is not chained yet)
        step("applyClearzone3", lambda: _lib.check(L.bbm_sitelist_clearzone3_dev(h, p(d_lists), p(d_nss), n, CAP, p(d_off), cp(pcfg), 0, p(d_out), None), "clearzone3"))
        # Read.setFromTopSite + SamLine(Read,int): FLAG / POS / MAPQ / RNAME of every read (no match strings yet: CIGAR '*')
        step("samFields", lambda: (_lib.check(L.bbm_sam_tasks_from_lists_dev(h, p(d_lists), p(d_nss), n, CAP, p(d_off), p(d_out), None, p(d_stasks), None), "sam tasks"),
                                   _lib.check(L.bbm_sam_batch_dev(h, p(d_stasks), n, p(d_nomatch), p(d_scaf_off), p(d_scaf_loc), p(d_scaf_len), len(co) - 1, cp(samcfg),
                                                                  p(d_souts), p(d_cigar), p(d_coff), None, None), "sam batch")))

    L.bbm_launch_count.restype = C.c_int64
    totals = []
    for rep in range(reps + 1):
        l0 = L.bbm_launch_count(h)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        chain()
        torch.cuda.synchronize(); totals.append(1e3 * (time.perf_counter() - t0))
        launches = L.bbm_launch_count(h) - l0
        if rep == 0:
            t.clear(); totals.clear()                                         # first pass is warm-up (allocations inside the library)
    lists = np.frombuffer(d_lists.cpu().numpy().tobytes(), sl.SS_DTYPE).reshape(n, CAP)
    nss = d_nss.cpu().numpy(); out = np.frombuffer(d_out.cpu().numpy().tobytes(), sl.READ_OUT_DTYPE); status = d_status.cpu().numpy()
    souts = np.frombuffer(d_souts.cpu().numpy().tobytes(), _sam.SAM_OUT_DTYPE)
    tr = R["truth"]; top = lists[:, 0]
    mapped = (out["flags"] & sl.F_MAPPED) != 0
    correct = mapped & (top["chrom"] == tr[:, 0]) & (top["strand"] == tr[:, 1]) & ((np.abs(top["start"] - tr[:, 2]) <= 8) | (np.abs(top["stop"] - tr[:, 3]) <= 8))
    exact = mapped & (top["start"] == tr[:, 2]) & (top["stop"] == tr[:, 3])
    med = float(np.median(totals))
    res = {"reference": "%d scaffold(s), %d chromosome array(s), %d index block(s)" % (scaffolds, len(co) - 1, idx.nblocks),
           "workload": "%d bp random reference, %d reads of 150 bp (the mates of %d pairs mapped as single reads), ~1%% subs, 1-3 bp indel in ~50%% of reads, Q30" % (genome_len, n, pairs),
           "reads": n, "ms": med, "reads_per_s": n / (med / 1e3), "launches_per_pass": int(launches),
           "stage_ms": {k: float(np.median(v)) for k, v in t.items()},
           "slow_alignments": int(na.value), "reads_slow_aligned": float(run_flags[0].float().mean().item()),
           "mapped": float(mapped.mean()), "top_site_is_origin": float(correct.mean()), "top_site_exact_start_and_stop": float(exact.mean()),
           "ambiguous": float(((out["flags"] & sl.F_AMBIGUOUS) != 0).mean()), "status_nonzero": int((status != 0).sum()), "status_gapped_site": int(((status & sl.SLOW_GAPPED) != 0).sum()),
           "status_aligner_error": int(((status & sl.SLOW_ALIGNER_ERROR) != 0).sum()),
           "mean_mapq": float(souts["mapq"][mapped].mean()) if mapped.any() else 0.0,
           "mean_sites_after_final": float(nss.mean()), "reads_lowered_by_clearzone3": float((out["best_sites"] > 0).mean()),
           "tip_deletion_sites_changed": int(np.frombuffer(d_out2.cpu().numpy().tobytes(), sl.READ_OUT_DTYPE)["best_sites"].sum()),
           "not_chained_yet": "genMatchString/realign_new (and the tip-score penalty and CIGAR that read its match string), pairing/rescue, SAM text",
           "timing": "host wall clock around device synchronisation, whole chain, median of %d passes after one warm-up pass" % reps}
    if cpu:
        cs = cpu_side(R, cb, co, table)
        m = cs["reads"]; cl = cs.pop("lists"); cn = cs.pop("nss"); cf = cs.pop("flags"); csam = cs.pop("sam")
        cs["device_sam_fields_identical_on_sample"] = bool(souts[:m].tobytes() == csam.tobytes())     # FLAG, POS, MAPQ, RNAME, RNEXT, PNEXT, TLEN, cigar length
        live = np.arange(CAP)[None, :] < cn[:, None]
        same = bool(np.array_equal(cn, nss[:m]) and np.array_equal(cf, out["flags"][:m]) and all(np.array_equal(cl[f][live], lists[:m][f][live]) for f in cl.dtype.names))
        cs["device_chain_identical_on_sample"] = same          # every field of every final site + the read flags, device chain vs CPU chain
        res["cpu_baseline"] = cs
    idx.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200_000)
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaffolds", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(run(a.pairs, a.genome, a.reps, cpu=not a.no_cpu, scaffolds=a.scaffolds)))


if __name__ == "__main__":
    main()
