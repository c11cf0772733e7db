"""GPU: bbm_map_batch_host (the whole chain behind one C-ABI call) == the sequential CPU chain (oracle/chain.py) on the same reads: every field of the
read record, the primary match string, FLAG/POS/MAPQ/CIGAR.  Unpaired reads; parity of the Java-only stages is UNPINNED against Java (no JVM)."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu

REC_FIELDS = ("chrom", "start", "stop", "strand", "map_score", "flags", "match_len", "cz3_sub", "tip_penalty", "status")


def _genome(kind, seed):
    g = wl.random_genome(300_000, seed=seed)
    if kind == "repeats":                       # planted repeat families: several sites per read, clearzone / ambiguity paths
        rng = np.random.Generator(np.random.PCG64(seed + 100))
        for _ in range(6):
            unit = wl.ACGT[rng.integers(0, 4, size=400, dtype=np.uint8)]
            for _c in range(5):
                u = unit.copy()
                m = rng.random(400) < 0.01
                u[m] = wl.ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
                q = int(rng.integers(0, len(g) - 400)); g[q:q + 400] = u
    return g


def _compare(dev, ref, n, flagged_ok=False):
    """flagged_ok: reads both sides flag with the same status bits (a match string longer than its slot: introns beyond ~1 kbp) are only required to agree
    on where they map; everything else must be identical field for field."""
    ms = ref["match_stride"]
    assert np.array_equal(dev["recs"]["status"], ref["recs"]["status"])
    clean = (ref["recs"]["status"] == 0) if flagged_ok else np.ones(n, bool)
    for f in REC_FIELDS:
        a, b = dev["recs"][f], ref["recs"][f]
        keep = clean | np.isin(f, ("chrom", "start", "stop", "strand", "status"))
        bad = np.nonzero((a != b) & keep)[0]
        assert len(bad) == 0, "record field %s differs for reads %s: device %s, oracle %s" % (f, bad[:5], a[bad[:5]], b[bad[:5]])
    dm = dev["match"][: n * ms].reshape(n, ms); om = ref["match"][: n * ms].reshape(n, ms)
    live = np.arange(ms)[None, :] < ref["recs"]["match_len"][:, None]
    assert np.array_equal(dm[live], om[live]), "primary match strings differ"
    for f in dev["sam"].dtype.names:
        a, b = dev["sam"][f], ref["sam"][f]
        bad = np.nonzero((a != b) & (clean if a.ndim == 1 else clean[:, None]))[0]
        assert len(bad) == 0, "SAM field %s differs for reads %s: device %s, oracle %s; records %s" % (f, bad[:5], a[bad[:5]], b[bad[:5]], ref["recs"][bad[:5]])


@pytest.mark.parametrize("kind,seed,L", [("plain", 11, 150), ("repeats", 12, 150), ("plain", 13, 100), ("repeats", 14, 250)])
def test_map_batch_single_equals_cpu_chain(kind, seed, L):
    from bbmap_b200.mapper import BBMapCUDA
    from oracle import chain, oracle as orc
    g = _genome(kind, seed)
    m = BBMapCUDA([g])
    try:
        R = wl.make_mapping_reads(m.cb, m.co, m.table, 1500, L=L, seed=seed + 1, sub_rate=0.015, indel_rate=0.02 / 3)
        o = orc.get()
        idx = o.index_build(m.cb, m.co, 13, -1)
        ref = chain.map_single(o, idx, m.cb, m.co, m.table, R["bases"], R["qual"], R["off"])
        dev = m.map_batch(R["bases"], R["qual"], R["off"], match_stride=ref["match_stride"])
        n = len(R["off"]) - 1
        assert ref["site_overflow"] == 0 and int(dev["stats"]["site_overflow_reads"]) == 0
        _compare(dev, ref, n)
        assert (ref["recs"]["flags"] & 1).mean() > 0.97
        assert int(dev["stats"]["realign_fills"]) == ref["realign_fills"] and int(dev["stats"]["slow_alignments"]) == ref["slow_alignments"]
    finally:
        m.close()


def test_map_batch_long_reads_broken_at_500():
    """configs[4]: 1-kbp single-ended reads cut by bbm_break_reads (`maxlen=500`, ReformatReads.breakReads) and mapped piece by piece: the device chain ==
    the sequential CPU chain on 500-row alignments (39 seeds per piece), names carry the piece number into the SAM text."""
    from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
    from bbmap_b200.reads import break_reads
    from oracle import chain, oracle as orc
    g = _genome("repeats", 31)
    m = BBMapCUDA([g])
    try:
        RL = wl.make_long_reads(m.cb, m.co, m.table, 300, L=1000, seed=32, sub_rate=0.015, indel_rate=0.0008)
        P = break_reads(RL["bases"], RL["qual"], RL["off"], RL["names"], RL["name_off"], 500, 0)
        n = len(P["src"])
        assert n == 600 and (np.diff(P["read_off"]) == 500).all()
        o = orc.get()
        idx = o.index_build(m.cb, m.co, 13, -1)
        ref = chain.map_single(o, idx, m.cb, m.co, m.table, P["bases"], P["quality"], P["read_off"])
        dev = m.map_batch(P["bases"], P["quality"], P["read_off"], cfg=mapper_cfg(sam_text=True), match_stride=ref["match_stride"], names=P["names"],
                          name_off=P["name_off"], sam_cap=n * 1400)
        assert ref["site_overflow"] == 0 and int(dev["stats"]["site_overflow_reads"]) == 0
        _compare(dev, ref, n)
        assert (ref["recs"]["flags"] & 1).mean() > 0.97
        assert int(dev["stats"]["realign_fills"]) == ref["realign_fills"] and int(dev["stats"]["slow_alignments"]) == ref["slow_alignments"]
        lines = bytes(dev["sam_text"][: int(dev["sam_off"][-1])]).split(b"\n")
        assert lines[0].startswith(b"long_0_1\t") and lines[1].startswith(b"long_0_2\t") and len(lines) >= n
    finally:
        m.close()


def test_map_batch_sam_text():
    """SAM lines (SamLine.toBytes + default tags) of the device chain == the Python restatement over the CPU chain's records; multi-scaffold reference."""
    from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
    from oracle import chain, oracle as orc
    scafs = [wl.random_genome(120_000, seed=21), wl.random_genome(90_000, seed=22), wl.random_genome(60_000, seed=23)]
    snames = ["chrA some description", "chrB", "scaffold_3"]
    m = BBMapCUDA(scafs, names=snames)
    try:
        R = wl.make_mapping_reads(m.cb, m.co, m.table, 1200, L=150, seed=24, sub_rate=0.02, indel_rate=0.02 / 3)
        n = len(R["off"]) - 1
        rng = np.random.Generator(np.random.PCG64(5))
        R["qual"] = rng.integers(2, 42, size=len(R["qual"])).astype(np.uint8)
        junk = wl.ACGT[rng.integers(0, 4, size=150 * 40, dtype=np.uint8)]                    # 40 unmappable reads at the end
        bases = np.concatenate([R["bases"], junk]); qual = np.concatenate([R["qual"], np.full(len(junk), 30, np.uint8)])
        off = np.arange(n + 41, dtype=np.int64) * 150
        names = [b"read_%d\tx/1" % i for i in range(n + 40)]
        nbuf = np.frombuffer(b"".join(names), np.int8).copy(); noff = np.zeros(n + 41, np.int64); np.cumsum([len(x) for x in names], out=noff[1:])
        o = orc.get()
        idx = o.index_build(m.cb, m.co, 13, -1)
        ref = chain.map_single(o, idx, m.cb, m.co, m.table, bases, qual, off)
        order = sorted(range(len(m.table)), key=lambda i: m.table[i])
        exp = chain.sam_lines(ref, off, names=names, scaf_names=[snames[i].encode() for i in order])
        dev = m.map_batch(bases, qual, off, cfg=mapper_cfg(sam_text=True), names=nbuf, name_off=noff, match_stride=ref["match_stride"], sam_cap=sum(len(x) for x in exp) + 64)
        _compare(dev, ref, n + 40)
        to = dev["sam_off"]
        assert int(to[-1]) == sum(len(x) for x in exp) == int(dev["stats"]["sam_bytes"])
        got = dev["sam_text"][: int(to[-1])].tobytes()
        assert got == b"".join(exp)
        assert (ref["recs"]["flags"][n:] & 1).sum() == 0 and (ref["recs"]["flags"][:n] & 1).mean() > 0.97
    finally:
        m.close()


@pytest.mark.parametrize("kind,seed,L,sub,indel", [("plain", 31, 150, 0.015, 0.02 / 3), ("repeats", 32, 150, 0.03, 0.03 / 3), ("plain", 33, 100, 0.05, 0.05 / 3),
                                                   ("repeats", 34, 250, 0.02, 0.02 / 3)])
def test_map_batch_pairs_equals_cpu_chain(kind, seed, L, sub, indel):
    """processReadPair: pairing, rescue, paired clearzone, genMatchString per mate, SAM pair fields — device chain == sequential CPU chain."""
    from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
    from oracle import chain, oracle as orc
    g = _genome(kind, seed)
    m = BBMapCUDA([g])
    try:
        R = wl.make_mapping_reads(m.cb, m.co, m.table, 1200, L=L, seed=seed + 1, sub_rate=sub, indel_rate=indel)
        n = len(R["off"]) - 1
        rng = np.random.Generator(np.random.PCG64(seed))
        bases = R["bases"].copy()
        for pidx in rng.choice(n // 2, size=60, replace=False):               # mates that only rescue can place: heavy damage on one side
            r = 2 * int(pidx) + int(rng.integers(0, 2)); a = r * L
            hit = rng.random(L) < 0.12
            bases[a:a + L][hit] = wl.ACGT[rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)]
        for pidx in rng.choice(n // 2, size=20, replace=False):               # and unmappable mates
            r = 2 * int(pidx) + 1; a = r * L
            bases[a:a + L] = wl.ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        o = orc.get()
        idx = o.index_build(m.cb, m.co, 13, -1)
        ref = chain.map_pairs(o, idx, m.cb, m.co, m.table, bases, R["qual"], R["off"])
        dev = m.map_batch(bases, R["qual"], R["off"], cfg=mapper_cfg(paired=True), match_stride=ref["match_stride"])
        assert ref["site_overflow"] == 0 and int(dev["stats"]["site_overflow_reads"]) == 0
        _compare(dev, ref, n)
        st = dev["stats"]
        assert (int(st["slow_alignments"]), int(st["realign_fills"]), int(st["rescue_scans"]), int(st["rescue_fills"]), int(st["mated_pairs"]), int(st["inner_length_sum"])) == \
               (ref["slow_alignments"], ref["realign_fills"], ref["rescue_scans"], ref["rescue_fills"], ref["mated"], ref["inner_sum"])
        f = ref["recs"]["flags"]
        assert ((f & 8) != 0).mean() > 0.9 and ((f & 16) != 0).sum() > 10          # most pairs mate; rescue placed some mates
    finally:
        m.close()


@pytest.mark.parametrize("paired", [True, False])
def test_map_batch_phix_fixture(paired):
    """configs[0]: the reference's own 100 shipped phiX pairs (tests/golden/phix.npz) through bbm_map_batch_host: device == CPU chain, and the mapping
    agrees with the origin the read names carry (the checks of tests/test_mapper_oracle.py)."""
    from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
    from oracle import chain, oracle as orc
    from test_mapper_oracle import _phix_pairs, check_mapping_invariants
    from test_search_oracle import phix
    d, _, _, _ = phix()
    cb, co, table, bases, qual, off, truth = _phix_pairs()
    m = BBMapCUDA([d["genome"]], names=["phiX"])
    try:
        assert np.array_equal(m.cb, cb)
        o = orc.get()
        idx = o.index_build(cb, co, 13, -1)
        ref = (chain.map_pairs if paired else chain.map_single)(o, idx, cb, co, table, bases, qual, off)
        dev = m.map_batch(bases, qual, off, cfg=mapper_cfg(paired=paired), match_stride=ref["match_stride"])
        _compare(dev, ref, len(off) - 1)
        dev["cigar"], dev["cigar_off"] = ref["cigar"], ref["cigar_off"]           # CIGAR text is a function of the (identical) records and match strings
        check_mapping_invariants(dev, off, truth)
    finally:
        m.close()


def test_map_batch_rescue_and_missing_mates():
    from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
    from oracle import chain, oracle as orc
    g = wl.random_genome(200_000, seed=41)
    m = BBMapCUDA([g])
    try:
        R = wl.make_mapping_reads(m.cb, m.co, m.table, 400, seed=42, sub_rate=0.01, indel_rate=0.0)
        L = 150
        rng = np.random.Generator(np.random.PCG64(43))
        bases = R["bases"].copy()
        for r in [2 * i + 1 for i in range(0, 60)]:
            v = bases[r * L:(r + 1) * L]
            v[4::8] = wl.ACGT[(np.searchsorted(wl.ACGT, v[4::8]) + 1) % 4]
        for r in [2 * i + 1 for i in range(100, 120)]:
            bases[r * L:(r + 1) * L] = wl.ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        o = orc.get()
        ref = chain.map_pairs(o, o.index_build(m.cb, m.co, 13, -1), m.cb, m.co, m.table, bases, R["qual"], R["off"])
        dev = m.map_batch(bases, R["qual"], R["off"], cfg=mapper_cfg(paired=True), match_stride=ref["match_stride"])
        _compare(dev, ref, len(R["off"]) - 1)
        assert ((ref["recs"]["flags"] & 16) != 0).sum() >= 40 and int(dev["stats"]["rescue_scans"]) == ref["rescue_scans"]
    finally:
        m.close()


def _edge_reads(scafs, rng):
    """Ragged and degenerate reads: empty, shorter than a k-mer, exactly one k-mer, long (600 = ALIGN_ROWS-1), all-N, N-rich, low quality, spliced over
    0.3-5 kbp introns (gapped sites; beyond ~1 kbp the match string outgrows its slot and both sides flag the read), reads across a scaffold boundary (removeOutOfBounds), reads from nowhere."""
    g = scafs[0]
    reads, quals = [], []
    def add(r, q=None):
        r = np.ascontiguousarray(r, np.uint8); reads.append(r)
        quals.append(np.full(len(r), 30, np.uint8) if q is None else np.ascontiguousarray(q, np.uint8))
    for L in (0, 5, 12, 13, 14, 30, 64, 100, 151, 300, 599, 600):
        p = int(rng.integers(1000, len(g) - 2000)); add(g[p:p + L])
        add(wl.revcomp(g[p + 700:p + 700 + L]))
    add(np.full(150, ord("N"), np.uint8)); add(np.full(40, ord("N"), np.uint8))
    for _ in range(12):                                                     # N-rich and low-quality reads
        p = int(rng.integers(1000, len(g) - 2000)); r = g[p:p + 150].copy()
        r[rng.random(150) < 0.15] = ord("N")
        q = rng.integers(2, 20, size=150).astype(np.uint8); q[r == ord("N")] = 0
        add(r, q)
    for _ in range(40):                                                     # spliced: 75 + intron + 75 (and 100 + 50)
        p = int(rng.integers(1000, len(g) - 8000)); gap = int(rng.integers(300, 1100)) if rng.random() < 0.7 else int(rng.integers(1100, 5000)); a = 75 if rng.random() < 0.5 else 100
        r = np.concatenate([g[p:p + a], g[p + a + gap:p + 150 + gap]])
        add(r if rng.random() < 0.5 else wl.revcomp(r))
    for k in range(6):                                                      # across the scaffold boundary of the packed chromosome (300 N in between)
        tail = scafs[0][len(scafs[0]) - 60 - 10 * k:]; head = scafs[1][:150 - len(tail)]
        add(np.concatenate([tail, head]))
    for _ in range(10):
        add(wl.ACGT[rng.integers(0, 4, size=150, dtype=np.uint8)])
    for _ in range(40):                                                     # ordinary reads with substitutions, so that the batch is not only corner cases
        p = int(rng.integers(1000, len(g) - 2000)); r = g[p:p + 150].copy()
        m_ = rng.random(150) < 0.03; r[m_] = wl.ACGT[rng.integers(0, 4, size=int(m_.sum()), dtype=np.uint8)]
        add(r if rng.random() < 0.5 else wl.revcomp(r))
    if len(reads) % 2:
        add(g[5000:5150])
    off = np.zeros(len(reads) + 1, np.int64); np.cumsum([len(r) for r in reads], out=off[1:])
    return np.concatenate(reads), np.concatenate(quals), off


@pytest.mark.parametrize("paired,slot", [(False, 0), (True, 0), (False, 6400), (True, 6400)])
def test_map_batch_edge_cases(paired, slot):
    """slot = 0: the default match-string slot (2 x longest read + 128): the spliced reads with introns beyond ~1 kbp are flagged MATCH_OVERFLOW by both sides.
    slot = 6400 (bbm_map_cfg.match_slot): every spliced read keeps its match string (up to 5 kbp of D) and must agree field for field, CIGAR included."""
    from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
    from oracle import chain, oracle as orc
    scafs = [wl.random_genome(150_000, seed=51), wl.random_genome(60_000, seed=52)]
    m = BBMapCUDA(scafs, names=["s1", "s2"])
    try:
        rng = np.random.Generator(np.random.PCG64(53))
        bases, qual, off = _edge_reads(scafs, rng)
        n = len(off) - 1
        o = orc.get()
        idx = o.index_build(m.cb, m.co, 13, -1)
        cfg = mapper_cfg(paired=paired, match_slot=slot)
        ref = (chain.map_pairs if paired else chain.map_single)(o, idx, m.cb, m.co, m.table, bases, qual, off, mcfg=cfg["map"].copy())
        dev = m.map_batch(bases, qual, off, cfg=cfg, match_stride=ref["match_stride"])
        _compare(dev, ref, n, flagged_ok=(slot == 0))
        f = ref["recs"]["flags"]; st = ref["recs"]["status"]
        if slot:
            assert (st != 0).sum() == 0 and ref["recs"]["match_len"].max() > 3000
        else:
            assert 0 < (st != 0).sum() < 30
        assert (f[:6] & 1).sum() == 0 and (f & 1).sum() > 100                 # reads shorter than a k-mer (0, 5, 12 bp) stay unmapped, the rest of the batch maps
        # degenerate batch sizes
        for cnt in (2, 0):
            sub = m.map_batch(bases[:off[cnt]], qual[:off[cnt]], off[:cnt + 1], cfg=cfg, match_stride=ref["match_stride"])
            assert len(sub["recs"]) == cnt
    finally:
        m.close()
