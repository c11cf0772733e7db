"""Per-stage device timings of the mapping core on BASELINE configs[1]-shaped input: E. coli-sized random
reference, 2x150 bp pairs with ~1% substitutions and 1-3 bp indels, Q=30.  Every stage runs through the C ABI with inputs
resident in HBM; times are the library's own CUDA-event measurements on its stream (kernel_ms_out).

    python bench/stages.py [--pairs 200000] [--genome 4600000] [--reps 5] [--stage all|ingest|seed|index|search|noindel|banded]

Prints one JSON object; bench.py embeds it as `stages`."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bbmap_b200 import lib as _lib, workloads as wl  # noqa: E402
from bbmap_b200.index import BBIndexCUDA, pack_chromosomes  # noqa: E402
from bbmap_b200.keyring import default_cfg  # noqa: E402
from bbmap_b200.search import HEAD_DTYPE, SITE_DTYPE  # noqa: E402

MAXK = 32          # slots per read for offsets/keys/keyScores (18 keys at 150 bp)
MAX_SITES = 16


def _nsites(L, h, block):
    n = C.c_int64(0)
    _lib.check(L.bbm_index_block_sites(h, block, C.byref(n)), "bbm_index_block_sites")
    return n.value


def cpu_side(R, cb, co, n_cpu=20000):
    """The same stages through the CPU oracle (C restatements, one host thread) on the first n_cpu reads — a reported baseline."""
    from oracle import oracle as orc
    o = orc.get()
    m = min(n_cpu, len(R["off"]) - 1)
    off = R["off"][:m + 1]; nb = int(off[-1])
    bases = R["bases"][:nb]; qual = R["qual"][:nb]
    out = {"reads": m, "cores": 1, "kind": "port"}
    t0 = time.perf_counter(); o.ingest_batch(bases, qual, off, 0); out["ingest_reads_per_s"] = m / (time.perf_counter() - t0)
    cfg = default_cfg()
    t0 = time.perf_counter(); seeds = o.seed_batch(bases, qual, off, cfg, MAXK); out["seed_reads_per_s"] = m / (time.perf_counter() - t0)
    t0 = time.perf_counter(); idx = o.index_build(cb, co, 13, -1); out["index_build_ms"] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); res = o.search_batch(idx, cb, co, bases, seeds["baseScores"], off, seeds, quit_after_two_perfects=False)
    out["search_reads_per_s"] = m / (time.perf_counter() - t0)
    ns = res["nsites"]; rid, sj = np.nonzero(np.arange(res["sites"].shape[1])[None, :] < ns[:, None])
    S = res["sites"][rid, sj]; S = S[S["strand"] == 0]; rid = rid[res["sites"][rid, sj]["strand"] == 0]
    tasks = np.zeros(len(S), wl.NOINDEL_TASK_DTYPE)
    tasks["read_off"] = off[rid]; tasks["ref_off"] = co[S["chrom"] - 1]; tasks["read_len"] = 150
    tasks["ref_len"] = (co[S["chrom"]] - co[S["chrom"] - 1]); tasks["ref_start"] = S["start"]
    t0 = time.perf_counter(); o.noindel_batch(bases, cb, tasks); out["noindel_sites_per_s"] = len(S) / (time.perf_counter() - t0)
    # quickRescue / findTipDeletions on the plus-strand sites (rescue searched with the read's own bases around its site: same scan cost)
    from bbmap_b200 import rescue as rs
    rt = np.zeros(len(S), rs.RESCUE_TASK_DTYPE)
    rt["read_off"] = tasks["read_off"]; rt["ref_off"] = tasks["ref_off"]; rt["read_len"] = 150; rt["ref_len"] = tasks["ref_len"]
    rt["max_index"] = rt["ref_len"] - 1; rt["loc"] = S["start"] - 400; rt["ideal_start"] = S["start"]; rt["search_dist"] = 600 + 250
    rt["max_mismatches"] = 32; rt["flags"] = 1
    t0 = time.perf_counter(); o.rescue_batch(bases, cb, rt, rs.rescue_cfg()); out["rescue_tasks_per_s"] = len(S) / (time.perf_counter() - t0)
    tt = np.zeros(len(S), rs.TIPDEL_TASK_DTYPE)
    tt["read_off"] = tasks["read_off"]; tt["ref_off"] = tasks["ref_off"]; tt["read_len"] = 150; tt["ref_len"] = tasks["ref_len"]
    tt["start"] = S["start"]; tt["stop"] = S["stop"]; tt["max_imperfect"] = 1; tt["flags"] = 3
    t0 = time.perf_counter(); o.tipdel_batch(bases, cb, tt, rs.tipdel_cfg()); out["tipdel_sites_per_s"] = len(S) / (time.perf_counter() - t0)
    return out


def run(pairs=200_000, genome_len=4_600_000, reps=5, stage="all", device=0, hbm_peak=6549.4, quit2=False, cpu=True, search_bps=0, scaffolds=1):
    import torch
    L = _lib.load()
    dev = torch.device("cuda", device)
    torch.cuda.set_device(device)
    if scaffolds <= 1:
        scafs = [wl.random_genome(genome_len, seed=1)]
    else:
        # G3: `scaffolds` random scaffolds of unequal size + 1 % of the bases in planted 300-bp repeat families of 100 copies
        rng = np.random.Generator(np.random.PCG64(3))
        w = rng.uniform(0.4, 2.0, size=scaffolds); sizes = np.maximum(1000, (w / w.sum() * genome_len).astype(np.int64))
        scafs = [wl.ACGT[rng.integers(0, 4, size=int(z), dtype=np.uint8)] for z in sizes]
        fam = max(1, int(genome_len * 0.01 / 300 / 100))
        for _ in range(fam):
            unit = wl.ACGT[rng.integers(0, 4, size=300, dtype=np.uint8)]
            for _c in range(100):
                sc = scafs[int(rng.integers(0, scaffolds))]
                p = int(rng.integers(0, len(sc) - 300)); sc[p:p + 300] = unit
    cb, co, table = pack_chromosomes(scafs)
    del scafs
    R = wl.make_mapping_reads(cb, co, table, pairs, seed=2)
    n = 2 * pairs
    nb = len(R["bases"])
    out = {"workload": "%d bp random reference (%d scaffold(s), %d chromosome array(s)%s), %d pairs 2x150, 1%% subs, 1-3 bp indel in ~50%% of reads, Q30"
                       % (genome_len, scaffolds, len(co) - 1, ", 1% planted repeats" if scaffolds > 1 else "", pairs),
           "reads": n}
    t0 = time.perf_counter()
    idx = BBIndexCUDA(cb, co, keylen=13, device=device)
    torch.cuda.synchronize()
    L.bbm_get_stat.restype = C.c_int64
    out["index_build"] = {"ms": L.bbm_get_stat(idx.h, b"index_build_us") / 1e3, "ms_with_upload_and_context": 1e3 * (time.perf_counter() - t0), "sites": int(sum(_nsites(L, idx.h, b) for b in range(idx.nblocks))), "blocks": idx.nblocks, "chrombits": int(idx.cfg["chrombits"][0]),
                          "note": "bbm_index_build with the reference resident: emit + radix sort + scan + analyzeIndex (COUNTS, clumpy keys, lengthHistogram), host-timed around a stream sync"}
    h = idx.h
    pad = lambda a, extra=64: torch.from_numpy(np.concatenate([a, np.zeros(extra, a.dtype)])).to(dev)
    d_bases = pad(R["bases"]); d_qual = pad(R["qual"]); d_off = torch.from_numpy(R["off"]).to(dev)
    d_basesM = torch.zeros(nb + 64, dtype=torch.uint8, device=dev); d_flags = torch.zeros(n, dtype=torch.int32, device=dev)
    ms = C.c_float(0)
    p = lambda t: C.c_void_p(t.data_ptr())

    def timed(fn):
        best = []
        for _ in range(reps):
            fn(); best.append(ms.value)
        return float(np.median(best))

    # ---- a0 ingest ----
    t = timed(lambda: _lib.check(L.bbm_ingest_batch_dev(h, p(d_bases), p(d_qual), p(d_off), n, 150, 0, p(d_basesM), p(d_flags), None, C.byref(ms)), "ingest"))
    alg = 5 * nb + 8 * n + 4 * n
    out["ingest"] = {"ms": t, "reads_per_s": n / (t / 1e3), "alg_bytes": alg, "GBps": alg / (t / 1e3) / 1e9, "frac_hbm": alg / (t / 1e3) / 1e9 / hbm_peak,
                     "bytes_per_read": alg / n}
    # ---- a1-a4 seed ----
    cfg = default_cfg()
    d_nkeys = torch.zeros(n, dtype=torch.int32, device=dev)
    d_offsets = torch.zeros(n * MAXK, dtype=torch.int32, device=dev); d_keys = torch.zeros_like(d_offsets); d_ks = torch.zeros_like(d_offsets)
    d_offM = torch.zeros_like(d_offsets); d_keysM = torch.zeros_like(d_offsets)
    d_bs = torch.zeros(nb + 64, dtype=torch.int8, device=dev)
    t = timed(lambda: _lib.check(L.bbm_seed_batch_dev(h, p(d_bases), p(d_qual), p(d_off), n, 150, cfg.ctypes.data_as(C.c_void_p), MAXK, p(d_nkeys), p(d_offsets),
                                                     p(d_keys), p(d_ks), p(d_bs), p(d_offM), p(d_keysM), None, C.byref(ms)), "seed"))
    nk = d_nkeys.cpu().numpy()
    alg = 2 * nb + nb + int(np.maximum(nk, 0).sum()) * 4 * 5 + 12 * n
    out["seed"] = {"ms": t, "reads_per_s": n / (t / 1e3), "alg_bytes": alg, "GBps": alg / (t / 1e3) / 1e9, "frac_hbm": alg / (t / 1e3) / 1e9 / hbm_peak,
                   "bytes_per_read": alg / n, "mean_keys": float(nk.mean())}
    # ---- a6-a9 search ----
    if os.environ.get("BBM_SEARCH_PROFILE"):
        _lib.check(L.bbm_set_option(h, b"search_profile", 1), "set_option")
    if os.environ.get("BBM_SEARCH_SPLIT"):
        _lib.check(L.bbm_set_option(h, b"search_split", int(os.environ["BBM_SEARCH_SPLIT"])), "set_option")
    if search_bps:
        _lib.check(L.bbm_set_option(h, b"search_shared", search_bps), "set_option")
    d_heads = torch.zeros(n * HEAD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_sites = torch.zeros(n * MAX_SITES * SITE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    t = timed(lambda: _lib.check(L.bbm_search_batch_dev(h, p(d_bases), p(d_bs), p(d_off), n, p(d_nkeys), p(d_offsets), p(d_ks), MAXK, 1 if quit2 else 0,
                                                       p(d_heads), p(d_sites), MAX_SITES, 150, None, C.byref(ms)), "search"))
    heads = np.frombuffer(d_heads.cpu().numpy().tobytes(), HEAD_DTYPE)
    sites = np.frombuffer(d_sites.cpu().numpy().tobytes(), SITE_DTYPE).reshape(n, MAX_SITES)
    # position-level truth: the top-scoring site is the read's origin
    ns = heads["nsites"]; tr = R["truth"]
    sc = np.where(np.arange(MAX_SITES)[None, :] < ns[:, None], sites["score"], -10 ** 9)
    bi = sc.argmax(axis=1); best = sites[np.arange(n), bi]
    correct = (ns > 0) & (best["chrom"] == tr[:, 0]) & (best["strand"] == tr[:, 1]) & ((np.abs(best["start"] - tr[:, 2]) <= 8) | (np.abs(best["stop"] - tr[:, 3]) <= 8))
    # algorithmic bytes: COUNTS gathers 4/key + per strand (starts pair 8 + list 4*len) for prescan and walk + ~2*(L+k) reference bytes per extended key
    counts = idx.download(0)[2] if idx.nblocks == 1 else None
    keys = d_keys.cpu().numpy().reshape(n, MAXK)
    valid = np.arange(MAXK)[None, :] < np.maximum(nk, 0)[:, None]
    if counts is None:       # several blocks: COUNTS is global, fetch it without the block's sites
        counts = np.zeros(1 << 26, np.int32)
        _lib.check(L.bbm_index_download(h, 0, None, None, counts.ctypes.data_as(C.c_void_p), None), "bbm_index_download")
    listlen = np.where(valid & (keys >= 0), counts[np.maximum(keys, 0)], 0).astype(np.int64)
    alg = int(valid.sum()) * 4 + 2 * (int(valid.sum()) * 2 * 8 + int(listlen.sum()) * 4) + int(valid.sum()) * (150 + 13) + n * (300 + 48) + int(ns.sum()) * 64
    out["search"] = {"ms": t, "reads_per_s": n / (t / 1e3), "alg_bytes": alg, "GBps": alg / (t / 1e3) / 1e9, "frac_hbm": alg / (t / 1e3) / 1e9 / hbm_peak,
                     "bytes_per_read": alg / n, "mean_sites": float(ns.mean()), "reads_with_site": float((ns > 0).mean()), "top_site_is_origin": float(correct.mean()),
                     "status_nonzero": int((heads["status"] != 0).sum())}
    if os.environ.get("BBM_SEARCH_PROFILE"):
        L.bbm_get_stat.restype = C.c_int64
        cyc = [L.bbm_get_stat(h, ("search_cycles_%d" % i).encode()) for i in range(5)]
        out["search"]["thread_cycle_shares"] = dict(zip(["filter", "prescan", "walk_incl_extend", "extend"], [round(x / max(cyc[0], 1), 3) for x in cyc[1:]]))
    # ---- a10 scoreNoIndels over every emitted site ----
    rid, sj = np.nonzero(np.arange(MAX_SITES)[None, :] < ns[:, None])
    S = sites[rid, sj]
    tasks = np.zeros(len(S), wl.NOINDEL_TASK_DTYPE)
    tasks["read_off"] = R["off"][rid]; tasks["ref_off"] = co[S["chrom"] - 1]; tasks["read_len"] = 150
    tasks["ref_len"] = (co[S["chrom"]] - co[S["chrom"] - 1]); tasks["ref_start"] = S["start"]
    plus = S["strand"] == 0
    d_tasks = torch.from_numpy(tasks.view(np.uint8)).to(dev); d_scores = torch.zeros(len(S), dtype=torch.int32, device=dev)
    d_chroms = C.c_void_p(idx.d_chroms.value)
    # minus-strand sites score the reverse complement: run the two strands against the matching read buffer
    ids = np.nonzero(plus)[0]
    tp = np.ascontiguousarray(tasks[ids]); d_tp = torch.from_numpy(tp.view(np.uint8)).to(dev)
    t1 = timed(lambda: _lib.check(L.bbm_noindel_batch_dev(h, p(d_bases), d_chroms, p(d_tp), p(d_scores), None, None, len(tp), None, C.byref(ms)), "noindel"))
    idm = np.nonzero(~plus)[0]
    tm = np.ascontiguousarray(tasks[idm]); d_tm = torch.from_numpy(tm.view(np.uint8)).to(dev)
    t2 = timed(lambda: _lib.check(L.bbm_noindel_batch_dev(h, p(d_basesM), d_chroms, p(d_tm), p(d_scores), None, None, len(tm), None, C.byref(ms)), "noindel"))
    t = t1 + t2
    alg = len(S) * (32 + 4 + 300)
    out["noindel"] = {"ms": t, "sites_per_s": len(S) / (t / 1e3), "alg_bytes": alg, "GBps": alg / (t / 1e3) / 1e9, "frac_hbm": alg / (t / 1e3) / 1e9 / hbm_peak}
    # ---- f3 mate rescue (quickRescue) and tip-deletion search over the top site of every read ----
    from bbmap_b200 import rescue as rs
    has = ns > 0
    anchors = np.nonzero(has)[0]
    A = best[anchors]; mate = anchors ^ 1
    rt = np.zeros(len(anchors), rs.RESCUE_TASK_DTYPE)
    into = (A["stop"] - A["start"] - 1 + 150 * 11 // 16).astype(np.int64)          # searchIntoAnchor (AbstractMapThread.java:1187)
    plusA = A["strand"] == 0
    rt["read_off"] = R["off"][mate]; rt["ref_off"] = co[A["chrom"] - 1]; rt["read_len"] = 150
    rt["ref_len"] = (co[A["chrom"]] - co[A["chrom"] - 1]); rt["min_index"] = 0; rt["max_index"] = rt["ref_len"] - 1
    rt["loc"] = np.where(plusA, A["stop"] - into, A["start"] + into)
    rt["ideal_start"] = np.where(plusA, A["stop"] + 100, A["start"] - 100)           # AVERAGE_PAIR_DIST=100
    rt["search_dist"] = 600 + into; rt["max_mismatches"] = 32; rt["flags"] = plusA.astype(np.int32)
    d_ro = torch.zeros(len(rt) * rs.RESCUE_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    rcfg = rs.rescue_cfg(); tot_ms = 0.0; found = 0; right_place = 0; starts_scanned = 0
    for strandA, buf in ((0, d_basesM), (1, d_bases)):                                # the loose mate is searched on the strand opposite to the anchor
        sel = np.nonzero(A["strand"] == strandA)[0]
        if not len(sel):
            continue
        tsel = np.ascontiguousarray(rt[sel]); d_rt = torch.from_numpy(tsel.view(np.uint8)).to(dev)
        tot_ms += timed(lambda: _lib.check(L.bbm_rescue_batch_dev(h, p(buf), d_chroms, p(d_rt), len(tsel), rcfg.ctypes.data_as(C.c_void_p), p(d_ro), None, C.byref(ms)), "rescue"))
        ro = np.frombuffer(d_ro[: len(tsel) * rs.RESCUE_OUT_DTYPE.itemsize].cpu().numpy().tobytes(), rs.RESCUE_OUT_DTYPE)
        found += int((ro["start"] >= 0).sum()); right_place += int((np.abs(ro["start"] - tr[mate[sel], 2]) <= 8).sum())
        starts_scanned += int((tsel["search_dist"].astype(np.int64) + 1).sum())
    alg = len(rt) * (56 + 32 + 150) + starts_scanned + len(rt) * 150
    out["rescue"] = {"ms": tot_ms, "tasks_per_s": len(rt) / (tot_ms / 1e3), "starts_per_s": starts_scanned / (tot_ms / 1e3), "found": found / max(1, len(rt)),
                     "found_at_mate_origin": right_place / max(1, len(rt)), "alg_bytes": alg, "GBps": alg / (tot_ms / 1e3) / 1e9,
                     "frac_hbm": alg / (tot_ms / 1e3) / 1e9 / hbm_peak,
                     "note": "quickRescue of each read's mate from the read's top site: searchDist 600+searchIntoAnchor, maxMismatches 32; algorithmic bytes = one pass over the search window + read + task/out records"}
    tt = np.zeros(len(S), rs.TIPDEL_TASK_DTYPE)
    tt["read_off"] = tasks["read_off"]; tt["ref_off"] = tasks["ref_off"]; tt["read_len"] = 150; tt["ref_len"] = tasks["ref_len"]
    tt["start"] = S["start"]; tt["stop"] = S["stop"]; tt["slow_score"] = 0; tt["max_imperfect"] = 1; tt["flags"] = 3
    d_to = torch.zeros(len(tt) * rs.TIPDEL_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    tcfg = rs.tipdel_cfg(); tot_ms = 0.0
    for sel, buf in ((ids, d_bases), (idm, d_basesM)):
        if not len(sel):
            continue
        tsel = np.ascontiguousarray(tt[sel]); d_tt = torch.from_numpy(tsel.view(np.uint8)).to(dev)
        tot_ms += timed(lambda: _lib.check(L.bbm_tipdel_batch_dev(h, p(buf), d_chroms, p(d_tt), len(tsel), tcfg.ctypes.data_as(C.c_void_p), p(d_to), None, C.byref(ms)), "tipdel"))
    alg = len(tt) * (48 + 16 + 2 * 8 + 2 * 8)
    out["tipdel"] = {"ms": tot_ms, "sites_per_s": len(tt) / (tot_ms / 1e3), "alg_bytes": alg, "GBps": alg / (tot_ms / 1e3) / 1e9, "frac_hbm": alg / (tot_ms / 1e3) / 1e9 / hbm_peak,
                     "note": "findTipDeletions (both tips) over every emitted site; most sites stop after the 8-base tip check (<3 mismatches)"}
    idx.close()
    # ---- a16 BandedAligner on G6-shaped pairs ----
    q, rf, bt = wl.make_banded_tasks(20_000, seed=6)
    from bbmap_b200.banded import BAND_OUT_DTYPE
    hb = C.c_void_p(); _lib.check(L.bbm_init(device, C.byref(hb)), "bbm_init")
    d_q = pad(q); d_r = pad(rf); d_bt = torch.from_numpy(bt.view(np.uint8)).to(dev)
    d_bo = torch.zeros(len(bt) * BAND_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    t = timed(lambda: _lib.check(L.bbm_banded_batch_dev(hb, p(d_q), p(d_r), p(d_bt), p(d_bo), len(bt), None, C.byref(ms)), "banded"))
    bo = np.frombuffer(d_bo.cpu().numpy().tobytes(), BAND_OUT_DTYPE)
    width = np.minimum(bt["max_width"], 2 * bt["max_edits"] + 1).astype(np.int64)
    cells = int((np.minimum(bt["query_len"], bt["ref_len"]).astype(np.int64) * width).sum())
    alg = int(bt["query_len"].sum() + bt["ref_len"].sum()) + len(bt) * (48 + 32)
    out["banded"] = {"ms": t, "pairs_per_s": len(bt) / (t / 1e3), "band_cells_upper_bound_gcups": cells / (t / 1e3) / 1e9, "alg_bytes": alg,
                     "GBps": alg / (t / 1e3) / 1e9, "frac_hbm": alg / (t / 1e3) / 1e9 / hbm_peak, "mean_edits": float(bo["edits"].mean())}
    L.bbm_destroy(hb)
    if cpu:
        out["cpu_baseline"] = cpu_side(R, cb, co)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200_000)
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--scaffolds", type=int, default=1)
    ap.add_argument("--search-bps", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    print(json.dumps(run(a.pairs, a.genome, a.reps, hbm_peak=hbm, search_bps=a.search_bps, cpu=not a.no_cpu, scaffolds=a.scaffolds)))


if __name__ == "__main__":
    main()
