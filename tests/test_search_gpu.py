"""GPU parity: BBIndex.find on the device (key filtering, prescan, slowWalk3, extendScore, site emission) vs the C
restatement — every field of every emitted site and the final bestScores[], bit-exact; plus the position-level truth check
the reference uses for synthetic reads (true origin must be the top site)."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _make(seed, sizes, n, L, k, paired=False, repeats=True, chrombits=-1, max_length=None):
    from bbmap_b200.index import pack_chromosomes
    rng = np.random.Generator(np.random.PCG64(seed))
    scafs = []
    for sz in sizes:
        s = wl.ACGT[rng.integers(0, 4, size=sz, dtype=np.uint8)].copy()
        if repeats and sz > 5000:
            unit = s[200:900].copy()
            for r in range(4):
                p = int(rng.integers(1000, sz - 800)); s[p:p + 700] = unit
                for q in rng.integers(0, 700, size=5):
                    s[p + q] = wl.ACGT[rng.integers(0, 4)]
            s[sz // 2: sz // 2 + 30] = ord("N")
        scafs.append(s)
    cb, co, table = pack_chromosomes(scafs, **({"max_length": max_length} if max_length else {}))
    reads, truth = [], []
    for i in range(n):
        si = int(rng.integers(0, len(scafs))); sc = scafs[si]
        p = int(rng.integers(0, len(sc) - L - 50))
        r = sc[p:p + L].copy()
        kind = rng.random()
        for q in rng.integers(0, L, size=int(rng.integers(0, 4))):
            r[q] = wl.ACGT[rng.integers(0, 4)]
        if kind < 0.15:      # small deletion / insertion
            q = int(rng.integers(20, L - 20)); d = int(rng.integers(1, 6))
            r = np.concatenate([sc[p:p + q], sc[p + q + d:p + L + d]]) if kind < 0.08 else np.concatenate([r[:q], wl.ACGT[rng.integers(0, 4, size=d, dtype=np.uint8)], r[q:L - d]])
        elif kind < 0.2:     # unrelated
            r = wl.ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        if rng.random() < 0.03:
            r[int(rng.integers(0, L))] = ord("N")
        st = int(rng.integers(0, 2))
        if st:
            r = wl.revcomp(r)
        reads.append(r); truth.append((table[si][0], st, table[si][1] + p, kind))
    bases = np.concatenate(reads)
    off = np.zeros(n + 1, np.int64); np.cumsum([len(r) for r in reads], out=off[1:])
    qual = rng.integers(20, 41, size=len(bases)).astype(np.uint8)
    qual[bases == ord("N")] = 0
    return cb, co, bases, qual, off, truth


# the last two have hit lists of ~6 and ~2 sites per key: they exercise the exact skip-ahead of the walks (long lists), one of them with two index blocks
@pytest.mark.parametrize("sizes,k,L,chrombits,maxlen", [((150000, 40000), 13, 150, -1, None), ((60000,) * 4, 11, 100, 1, 90000), ((30000,), 10, 250, 0, None),
                                                        ((1500000,), 9, 150, -1, None), ((700000,) * 3, 10, 100, 1, 1000000)])
@pytest.mark.parametrize("quit2", [True, False])
def test_search_parity(oracle, sizes, k, L, chrombits, maxlen, quit2):
    from bbmap_b200.index import BBIndexCUDA
    from bbmap_b200.keyring import KeyRingCUDA, default_cfg
    from bbmap_b200 import search
    cb, co, bases, qual, off, truth = _make(7 + k, sizes, 1500, L, k, chrombits=chrombits, max_length=maxlen)
    cfg = default_cfg(); cfg["keylen"] = k; cfg["baseKeyHitScore"] = 100 * k
    eidx = oracle.index_build(cb, co, k, chrombits)
    eseeds = oracle.seed_batch(bases, qual, off, cfg, 96)
    exp = oracle.search_batch(eidx, cb, co, bases, eseeds["baseScores"], off, eseeds, quit_after_two_perfects=quit2)
    idx = BBIndexCUDA(cb, co, keylen=k, chrombits=chrombits)
    try:
        kr = KeyRingCUDA(ctx=idx.h)
        seeds = kr.seed_batch(bases, qual, off, cfg, 96)
        for key in ("nkeys", "offsets", "keyScores", "baseScores"):
            assert np.array_equal(seeds[key], eseeds[key])
        heads, sites = search.search_batch(idx, bases, seeds["baseScores"], off, seeds, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2)
        # the same batch with one phase of BBIndex.find per kernel launch (key filtering / prescan / walk)
        h3, t3 = search.search_batch(idx, bases, seeds["baseScores"], off, seeds, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2, split=True)
        assert h3.tobytes() == heads.tobytes() and t3.tobytes() == sites.tobytes()
        # ... and as a single thread-per-read launch (the default above is the split with the warp-per-read prescan; reads with more than
        # 32 keys fall back to the thread-per-read prescan)
        h4, t4 = search.search_batch(idx, bases, seeds["baseScores"], off, seeds, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2, split=0)
        assert h4.tobytes() == heads.tobytes() and t4.tobytes() == sites.tobytes()
        # ... and with the warp-per-read prescan but the thread-per-read walk (the default, split=3, walks with one warp per read as well)
        h5, t5 = search.search_batch(idx, bases, seeds["baseScores"], off, seeds, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2, split=2)
        assert h5.tobytes() == heads.tobytes() and t5.tobytes() == sites.tobytes()
        # the same batch with 32 key slots per read through the shared-memory variant of the kernel
        s32 = {k: (np.ascontiguousarray(v[:, :32]) if getattr(v, "ndim", 1) == 2 else v) for k, v in seeds.items()}
        if int(seeds["nkeys"].max()) <= 32:
            h2, t2 = search.search_batch(idx, bases, seeds["baseScores"], off, s32, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2, shared=True)
            assert h2.tobytes() == heads.tobytes() and t2.tobytes() == sites.tobytes()
            # 32 key slots: no thread-per-read launch follows the warp-per-read walk at all
            h6, t6 = search.search_batch(idx, bases, seeds["baseScores"], off, s32, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2)
            assert h6.tobytes() == heads.tobytes() and t6.tobytes() == sites.tobytes()
    finally:
        idx.close()
    for f in ("nsites", "status", "num_hits", "max_score", "max_quick_score", "best_scores"):
        assert np.array_equal(heads[f], exp[f]), f
    assert ((exp["status"] & ~4) == 0).all()      # only "gap array longer than 9 ints" may be flagged (tiny genome, k=10)
    for i in range(len(heads)):
        ns = exp["nsites"][i]
        assert sites[i, :ns].tobytes() == exp["sites"][i, :ns].tobytes(), (i, sites[i, :ns], exp["sites"][i, :ns])
    # position-level truth (AbstractMapThread.isCorrectHit-style): related reads map to their origin with the top score
    good = tot = 0
    for i, (chrom, st, pos, kind) in enumerate(truth):
        if kind >= 0.15 and kind < 0.2:
            continue
        tot += 1
        ns = heads["nsites"][i]
        if ns == 0:
            continue
        s = sites[i, :ns]; best = s[np.argmax(s["score"])]
        good += int(best["chrom"] == chrom and best["strand"] == st and abs(int(best["start"]) - pos) <= 8)
    assert good >= 0.95 * tot, (good, tot)
    assert (heads["nsites"] > 1).any()


@pytest.mark.parametrize("tag", ["r1", "r2"])
def test_search_phix_fixture(oracle, tag):
    """The reference's own shipped test reads (configs[0]): device == oracle bit for bit, and the true origin (in the read
    name) is the top site for >= 95 of 100 reads."""
    from test_search_oracle import phix, top_site_correct
    from bbmap_b200.index import BBIndexCUDA
    from bbmap_b200.keyring import KeyRingCUDA, default_cfg
    from bbmap_b200 import search
    d, cb, co, table = phix()
    b, q, off, truth = d[tag + "_bases"], d[tag + "_qual"], d[tag + "_off"], d[tag + "_truth"]
    eidx = oracle.index_build(cb, co, 13, -1)
    eseeds = oracle.seed_batch(b, q, off, default_cfg(), 96)
    idx = BBIndexCUDA(cb, co, keylen=13)
    try:
        seeds = KeyRingCUDA(ctx=idx.h).seed_batch(b, q, off, default_cfg(), 96)
        for quit2 in (True, False):
            exp = oracle.search_batch(eidx, cb, co, b, eseeds["baseScores"], off, eseeds, quit_after_two_perfects=quit2)
            heads, sites = search.search_batch(idx, b, seeds["baseScores"], off, seeds, quit_after_two_perfects=quit2)
            for f in ("nsites", "status", "num_hits", "max_score", "max_quick_score", "best_scores"):
                assert np.array_equal(heads[f], exp[f]), f
            for i in range(len(heads)):
                ns = exp["nsites"][i]
                assert sites[i, :ns].tobytes() == exp["sites"][i, :ns].tobytes()
            assert sum(top_site_correct(sites[i], heads["nsites"][i], truth[i]) for i in range(len(truth))) >= 95
    finally:
        idx.close()


@pytest.mark.parametrize("L,seed", [(250, 14), (150, 31), (100, 32)])
def test_search_parity_repeat_families(oracle, L, seed):
    """Near-identical 400-bp repeat families (five copies each, 1 % apart) and reads with indels: extensions whose keys point at several copies,
    so extendScore's per-key loops stop at every kind of cell — the case that caught a chunk-boundary slip in the warp-per-read walk (a stop
    after the 32nd position of a chunk).  Every launch variant against the C restatement."""
    from bbmap_b200.index import BBIndexCUDA, pack_chromosomes
    from bbmap_b200.keyring import default_cfg
    from bbmap_b200 import search
    g = wl.random_genome(300_000, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    starts = []
    for _ in range(8):
        unit = wl.ACGT[rng.integers(0, 4, size=400, dtype=np.uint8)]
        for _c in range(5):
            u = unit.copy()
            m = rng.random(400) < 0.01
            u[m] = wl.ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
            q = int(rng.integers(0, len(g) - 400)); g[q:q + 400] = u; starts.append(q)
    cb, co, table = pack_chromosomes([g])
    R = wl.make_mapping_reads(cb, co, table, 1500, L=L, seed=seed + 1, sub_rate=0.015, indel_rate=0.02 / 3)
    bases, qual, off = R["bases"].copy(), R["qual"], R["off"]
    for i in range(0, len(off) - 1, 2):                       # every second read from inside a repeat copy (clean), either strand
        q = starts[int(rng.integers(0, len(starts)))] + int(rng.integers(-L // 2, 400 - L // 2))
        q = min(max(q, 0), len(g) - L)
        r = g[q:q + L]
        bases[off[i]:off[i + 1]] = r if (i // 2) % 2 == 0 else wl.revcomp(r)
    cfg = default_cfg()
    eidx = oracle.index_build(cb, co, 13, -1)
    es = oracle.seed_batch(bases, qual, off, cfg, 32)
    idx = BBIndexCUDA(cb, co, keylen=13)
    try:
        for quit2 in (True, False):
            exp = oracle.search_batch(eidx, cb, co, bases, es["baseScores"], off, es, quit_after_two_perfects=quit2)
            for split in (3, 2):
                h, t = search.search_batch(idx, bases, es["baseScores"], off, es, max_sites=search.MAX_SITES, quit_after_two_perfects=quit2, split=split)
                for f in ("nsites", "status", "num_hits", "max_score", "max_quick_score", "best_scores"):
                    assert np.array_equal(h[f], exp[f]), (f, split, quit2, np.nonzero((h[f] != exp[f]).reshape(len(h), -1).any(axis=1))[0][:5])
                for i in range(len(h)):
                    ns = exp["nsites"][i]
                    assert t[i, :ns].tobytes() == exp["sites"][i, :ns].tobytes(), (i, split, quit2)
            assert (exp["nsites"] > 2).sum() > 200
    finally:
        idx.close()
