"""TEST INFRASTRUCTURE ONLY — the whole mapping chain through the CPU oracle (C restatements in liborc.so, the reference's own C for the fills
when oracle/_ref exists is NOT used here: the chain calls the port, which tests/test_oracle_vs_reference.py pins to the reference's C).

    Read.validate -> quickMap (KeyRing, BBIndex.find, removeOutOfBounds) -> trimList -> scoreNoIndels -> findTipDeletions -> scoreSlow ->
    final list policy -> genMatchString / realign_new -> clearzone 3 / toLocalAlignment / score gates / tip penalty -> SamLine fields

Used by tests/ (the checker of bbm_map_batch_*) and by bench.py's cpu_baseline / --impl reference legs."""
import numpy as np

from bbmap_b200 import mapper as mp, sitelist as sl
from bbmap_b200.keyring import default_cfg
from bbmap_b200.rescue import tipdel_cfg
from bbmap_b200.sam import scaffold_table, default_cfg as sam_default_cfg, SAM_TASK_DTYPE

MAXK = 32


def max_keys_for(off):
    """key slots per read: 32 covers reads up to ~200 bp; longer reads get 2 keys per k bases + 3 (the key density never exceeds 1.9, AbstractMapThread.java:663-676), at most 96"""
    longest = int(np.diff(off).max()) if len(off) > 1 else 0
    return max(MAXK, min(96, 2 * longest // 13 + 3))


def match_stride(max_len, mcfg=None):
    slot = int(mcfg["match_slot"][0]) if mcfg is not None and "match_slot" in mcfg.dtype.names else 0
    return max(2 * int(max_len) + 128, slot)


def search_to_lists(res, cap):
    m = len(res["nsites"])
    ns = np.minimum(res["nsites"], cap).astype(np.int32)
    lists = np.zeros((m, cap), sl.SS_DTYPE); S = res["sites"][:, :cap]
    for f in ("chrom", "start", "stop", "hits", "score", "strand", "perfect", "semiperfect", "ngaps", "gaps"):
        lists[f] = S[f]
    lists["quick_score"] = S["score"]
    return lists, ns


def map_single(o, idx, cb, co, table, bases, qual, off, cap=16, pcfg=None, mcfg=None, sam_cfg=None, ingest_flags=0):
    """Unpaired reads.  idx = o.index_build(cb, co, 13, -1).  Returns a dict with the final lists, the read records, match strings, SAM
    fields and CIGAR text."""
    pcfg = sl.policy_cfg() if pcfg is None else pcfg
    mcfg = mp.map_cfg() if mcfg is None else mcfg
    scfg = sam_default_cfg() if sam_cfg is None else sam_cfg
    off = np.ascontiguousarray(off, np.int64); n = len(off) - 1
    nb = int(off[-1])
    bases, qual, basesM, rflags = o.ingest_batch(bases[:nb], None if qual is None else qual[:nb], off, ingest_flags)
    seeds = o.seed_batch(bases, qual, off, default_cfg(), max_keys_for(off))
    res = o.search_batch(idx, cb, co, bases, seeds["baseScores"], off, seeds, quit_after_two_perfects=True)     # AbstractIndex.QUIT_AFTER_TWO_PERFECTS, single-ended
    overflow = int(((res["status"] & 2) != 0).sum() + (res["nsites"] > cap).sum())
    lists, ns = search_to_lists(res, cap)
    scaf = scaffold_table(table, len(co) - 1)
    maxidx = (np.diff(np.asarray(co, np.int64)) - 1).astype(np.int32)
    lists, ns, _ = o.sitelist_bounds(lists, ns, off, maxidx, scaf)
    lists, ns, _ = o.sitelist(sl.SL_TRIM, lists, ns, off, pcfg)
    lists, ns, out = o.sitelist(sl.SL_NOINDEL, lists, ns, off, pcfg, bases, basesM, cb, co)
    runm = (out["near_perfect"] < 1).astype(np.int32)
    lists, _ = o.sitelist_tipdel(lists, ns * runm, off, bases, basesM, qual, cb, co, tipdel_cfg())
    lists, slow_status, na = o.score_slow(lists, ns, off, bases, basesM, cb, co, runm, sl.slow_cfg())
    lists, ns, out = o.sitelist(sl.SL_FINAL, lists, ns, off, pcfg)
    ms = match_stride(int(np.diff(off).max()) if n else 1, mcfg)
    lists, ns, recs, match, fills = o.map_finish_single(lists, ns, off, bases, basesM, cb, co, out, pcfg, mcfg, ms)
    recs["flags"] |= np.where(seeds["nkeys"] < 0, 32, 0).astype(np.int32)          # quickMap < 0: r.setDiscarded(true)
    tasks = sam_tasks(recs, off, ms)
    srec, cig, cig_off = o.sam_batch(tasks, match, scaf, scfg)
    return {"lists": lists, "nss": ns, "recs": recs, "match": match, "match_stride": ms, "sam": srec, "cigar": cig, "cigar_off": cig_off,
            "slow_alignments": int(na), "realign_fills": int(fills), "site_overflow": overflow, "discarded": (seeds["nkeys"] < 0), "bases": bases, "basesM": basesM,
            "qual": qual}


def map_pairs(o, idx, cb, co, table, bases, qual, off, cap=16, pcfg=None, mcfg=None, sam_cfg=None, ingest_flags=0):
    """Paired reads (reads 2i / 2i+1 are mates): BBMapThread.processReadPair."""
    pcfg = sl.policy_cfg() if pcfg is None else pcfg
    mcfg = mp.map_cfg(paired=1) if mcfg is None else mcfg
    scfg = sam_default_cfg() if sam_cfg is None else sam_cfg
    off = np.ascontiguousarray(off, np.int64); n = len(off) - 1
    nb = int(off[-1])
    bases, qual, basesM, rflags = o.ingest_batch(bases[:nb], None if qual is None else qual[:nb], off, ingest_flags)
    seeds = o.seed_batch(bases, qual, off, default_cfg(), max_keys_for(off))
    res = o.search_batch(idx, cb, co, bases, seeds["baseScores"], off, seeds, quit_after_two_perfects=False)    # forced false in paired mode (BBMap.java:434)
    overflow = int(((res["status"] & 2) != 0).sum() + (res["nsites"] > cap).sum())
    lists, ns = search_to_lists(res, cap)
    scaf = scaffold_table(table, len(co) - 1)
    maxidx = (np.diff(np.asarray(co, np.int64)) - 1).astype(np.int32)
    lists, ns, _ = o.sitelist_bounds(lists, ns, off, maxidx, scaf)
    ms = match_stride(int(np.diff(off).max()) if n else 1, mcfg)
    wcfg = sl.slow_cfg(paired=1, min_ratio=mcfg["min_ratio"][0], min_ratio_pre_rescue=mcfg["min_ratio_pre_rescue"][0])
    lists, ns, recs, match, stats = o.map_pairs(lists, ns, off, bases, basesM, qual, cb, co, seeds["nkeys"], pcfg, mcfg, wcfg, tipdel_cfg(), ms)
    tasks = sam_tasks(recs, off, ms, mate=np.arange(n, dtype=np.int32) ^ 1)
    tasks["flags"] |= np.where(np.arange(n) & 1, 128, 0).astype(np.int32)
    srec, cig, cig_off = o.sam_batch(tasks, match, scaf, scfg)
    return {"lists": lists, "nss": ns, "recs": recs, "match": match, "match_stride": ms, "sam": srec, "cigar": cig, "cigar_off": cig_off,
            "slow_alignments": int(stats[0]), "realign_fills": int(stats[1]), "rescue_scans": int(stats[2]), "rescue_fills": int(stats[3]), "mated": int(stats[4]),
            "inner_sum": int(stats[5]), "site_overflow": overflow, "bases": bases, "basesM": basesM, "qual": qual}


def sam_tasks(recs, off, ms, mate=None):
    """Read fields -> the record SamLine(Read,int) reads (bbm_sam_task)."""
    n = len(recs)
    t = np.zeros(n, SAM_TASK_DTYPE)
    t["match_off"] = np.arange(n, dtype=np.int64) * ms
    t["match_len"] = recs["match_len"]
    t["chrom"] = recs["chrom"]; t["start"] = recs["start"]; t["stop"] = recs["stop"]
    t["read_len"] = np.diff(off).astype(np.int32)
    t["score"] = recs["map_score"]
    t["mate"] = -1 if mate is None else mate
    f = recs["flags"]
    t["flags"] = ((f & 1) * 1) | (np.where((recs["strand"] == 1) & ((f & 1) != 0), 2, 0)) | (np.where(f & 2, 4, 0)) | (np.where(f & 4, 8, 0)) | (np.where(f & 32, 32, 0)) | (np.where(f & 8, 64, 0))
    return t


def sam_lines(res, off, names=None, scaf_names=None, paired=False, intron_limit=2 ** 31 - 1):
    """SamLine.toBytes (current/stream/SamLine.java:1925-1960) + makeOptionalTags (:1481-1560, the default NM / AM tags and XT:A:R) for every read of a
    chain result, in plain Python: QNAME rule of SamLine(Read,int) (:100-112), reverse-complemented SEQ and reversed QUAL for mapped minus-strand lines."""
    recs, sam, ms = res["recs"], res["sam"], res["match_stride"]
    bases, basesM, qual = res["bases"], res["basesM"], res["qual"]
    lines = []
    for r in range(len(recs)):
        q, o = recs[r], sam[r]
        a, b = int(off[r]), int(off[r + 1])
        if names is None:
            qname = b"*"
        else:
            qname = names[r].replace(b"\t", b"_")
            if paired and len(qname) > 2 and qname[-1:] in (b"1", b"2") and qname[-2:-1] in (b" ", b"/"):
                qname = qname[:-2]
        def sname(i):
            return b"*" if (i < 0 or scaf_names is None) else scaf_names[i]
        cig = b"*" if o["cigar_len"] <= 0 else res["cigar"][int(res["cigar_off"][r]):int(res["cigar_off"][r]) + int(o["cigar_len"])].tobytes()
        rnext = b"*" if o["rnext"] == -1 else (b"=" if o["rnext"] == -2 else sname(int(o["rnext"])))
        mapped_line = (o["flag"] & 4) == 0; minus = (o["flag"] & 16) != 0
        L = b - a
        if L == 0:
            seq = b"*"; ql = b"*"
        else:
            rc = mapped_line and minus
            seq = (basesM if rc else bases)[a:b].tobytes()
            if qual is None:
                ql = b"*"
            else:
                qq = qual[a:b].astype(np.int16) + 33
                ql = bytes((qq[::-1] if rc else qq).astype(np.uint8).tolist())
        f = [qname, b"%d" % o["flag"], sname(int(o["scaffold"])), b"%d" % o["pos"], b"%d" % o["mapq"], cig, rnext, b"%d" % o["pnext"], b"%d" % o["tlen"], seq, ql]
        if q["flags"] & 1:
            if q["flags"] & 4:
                f.append(b"XT:A:R")
            if q["flags"] & 2:
                f.append(b"NM:i:0")
            elif q["match_len"] > 0:
                m = res["match"][r * ms:r * ms + int(q["match_len"])].tobytes()
                import re
                c = cig.decode() if cig != b"*" else ""
                lm = re.match(r"^(\d+)S", c); rm = re.search(r"(\d+)S$", c)
                lo = int(lm.group(1)) if lm else 0; hi = L - (int(rm.group(1)) if rm else 0)
                nm = 0; dels = 0; cpos = 0
                for ch in m:
                    if lo <= cpos < hi:
                        if ch in b"ISNXY":
                            nm += 1
                        if ch == ord("D"):
                            dels += 1
                        else:
                            if dels <= intron_limit:
                                nm += dels
                            dels = 0
                    if ch != ord("D"):
                        cpos += 1
                if dels <= intron_limit:
                    nm += dels
                f.append(b"NM:i:%d" % nm)
            am = int(o["mapq"])
            if paired:
                mt = recs[r ^ 1]; ml = int(off[(r ^ 1) + 1] - off[r ^ 1])
                am = min(am, (max(1, int(mt["map_score"]) // ml) if (mt["flags"] & 1) else 0))
            f.append(b"AM:i:%d" % am)
        lines.append(b"\t".join(f) + b"\n")
    return lines
