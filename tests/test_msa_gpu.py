"""GPU parity: the CUDA MultiStateAligner11ts path (through the C ABI) vs the CPU oracle, bit-exact:
result vector, which fill ran, the reference's iteration counter, score2 vector, match string."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from kat import KATS, REF, MINSCORE, B, rle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["strip", "tiled", "mixed"])
def msa(request):
    """Both routings of limited un-banded fills: the thread-per-alignment strip kernel (default) and the register-tiled kernel."""
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    m = MultiStateAligner11tsCUDA()
    m.set_option("strip", {"strip": 16, "tiled": 0, "mixed": 3}[request.param])
    m.set_option("strip_min_tasks", 0)     # the library keeps small batches away from the thread-per-alignment kernel; the tests want it exercised
    if request.param != "strip":
        m.set_option("narrow", 1)          # every shape-eligible alignment tries the narrow kernel first: exercises the hand-over path
    if request.param == "tiled":
        m.set_option("band", 0)            # banded fills through the register-tiled kernel + row-sequential re-runs (the other two routings use msa_band.cu)
    yield m
    m.close()


def _compare(oracle, msa, reads, genome, tasks, bw=0, ratio=0.0, kind=None):
    """kind=None: the fills of the comparator are the reference's OWN C (oracle/_ref/libbbref.so, prebuilt, travels to the GPU box)
    whenever it is there; the C restatement ("port") otherwise.  port == reference is a CPU test (tests/test_oracle_vs_reference.py)."""
    if kind is None:
        kind = "reference" if oracle.has_reference else "port"
    moff = wl.match_offsets(tasks)
    exp, emb, cells = oracle.run_batch(reads, genome, tasks, match_off=moff, bandwidth=bw, ratio=ratio, kind=kind, threads=8)
    msa.set_band(bw, ratio)
    d_ref = msa.load_reference(genome)
    try:
        got, gmb = msa.align_batch(reads, d_ref, tasks, match_off=moff)
    finally:
        msa.free(d_ref)
    bad = [i for i in range(len(tasks)) if got[i].tobytes() != exp[i].tobytes()]
    if bad:
        i = bad[0]
        raise AssertionError("%d/%d tasks differ; first %d\n task=%s\n got=%s\n exp=%s" % (len(bad), len(tasks), i, tasks[i], got[i], exp[i]))
    for i in range(len(tasks)):
        n = exp[i]["match_len"]
        if n > 0:
            a = moff[i]
            assert gmb[a:a + n].tobytes() == emb[a:a + n].tobytes(), "match string differs for task %d: %s vs %s" % (
                i, rle(gmb[a:a + n]), rle(emb[a:a + n]))
    return exp


def test_kats(oracle, msa):
    ref = B(REF)
    for name, read, a, b, fn, bw, result, iters, score2, match in KATS:
        r = B(read)
        tasks = np.zeros(1, wl.TASK_DTYPE)
        tasks[0] = (0, 0, len(r), len(ref), a, b, MINSCORE, (wl.TF_RAW_LIMITED if fn == "limited" else wl.TF_RAW_UNLIMITED) | wl.TF_SCORE | wl.TF_TRACEBACK)
        msa.set_band(bw, 0.0)
        d_ref = msa.load_reference(ref)
        outs, mbuf = msa.align_batch(r, d_ref, tasks)
        msa.free(d_ref)
        o = outs[0]
        assert o["status"] == 0, name
        assert o["result"][: len(result)].tolist() == result, name
        if iters is not None:
            assert o["iterations"] == iters, name
        if score2 is None:
            assert o["score_len"] == 0 and o["match_len"] == -1
        else:
            assert o["score"][: len(score2)].tolist() == score2 and o["score_len"] == len(score2), name
            assert rle(mbuf[: o["match_len"]]) == match, name


@pytest.mark.parametrize("mode", ["java", "raw_limited", "raw_unlimited"])
@pytest.mark.parametrize("tight", [True, False])
def test_random_noband(oracle, msa, mode, tight):
    genome = wl.random_genome(50000, seed=21)
    flags = wl.TF_SCORE | wl.TF_TRACEBACK | {"java": 0, "raw_limited": wl.TF_RAW_LIMITED, "raw_unlimited": wl.TF_RAW_UNLIMITED}[mode]
    reads, tasks = wl.make_msa_tasks(genome, 3000 if mode != "raw_unlimited" else 600, seed=31, flags=flags, tight=tight,
                                     ratio=0.56 if tight else 0.336)
    if mode == "raw_limited":
        tasks["min_score"] -= 120
    exp = _compare(oracle, msa, reads, genome, tasks)
    if mode != "raw_unlimited":
        assert (exp["result"][:, 4] == 1).any() and (exp["match_len"] > 0).any()


@pytest.mark.parametrize("bw,ratio", [(12, 0.0), (40, 0.0), (0, 0.18), (8, 0.3)])
def test_random_banded(oracle, msa, bw, ratio):
    genome = wl.random_genome(50000, seed=22)
    for tight in (True, False):
        reads, tasks = wl.make_msa_tasks(genome, 1500, seed=41 + bw, flags=wl.TF_SCORE | wl.TF_TRACEBACK, tight=tight)
        _compare(oracle, msa, reads, genome, tasks, bw=bw, ratio=ratio)


def test_odd_shapes(oracle, msa):
    """short reads (unlimited rule), long reads, wide windows (generic kernel), N-rich reference, clamped windows."""
    genome = wl.random_genome(30000, seed=23).copy()
    genome[5000:5040] = ord("N")
    genome[:50] = ord("N")
    genome[-50:] = ord("N")
    rng = np.random.Generator(np.random.PCG64(7))
    reads_l, tasks_l, off = [], [], 0
    def add(pos, L, a, b, ms, flags):
        nonlocal off
        r = genome[pos:pos + L].copy()
        for k in rng.integers(0, L, size=max(1, L // 60)):
            r[k] = wl.ACGT[rng.integers(0, 4)]
        reads_l.append(r)
        tasks_l.append((off, 0, L, len(genome), a, b, ms, flags))
        off += L
    F = wl.TF_SCORE | wl.TF_TRACEBACK
    for L in (20, 31, 40, 45, 64, 100, 199, 250, 300, 400, 600):
        for pad in (0, 4, 17, 60):
            pos = int(rng.integers(200, len(genome) - L - 300))
            add(pos, L, pos - pad, pos + L - 1 + pad, int(0.5 * wl.max_quality(L)), F)
            add(pos, L, pos - pad, pos + L - 1 + pad, int(0.5 * wl.max_quality(L)), F | wl.TF_RAW_LIMITED)
            add(pos, L, pos - pad, pos + L - 1 + pad, 0, F | wl.TF_RAW_UNLIMITED)
    # windows over the N block, over the array ends (clamped), and narrower than the read
    for L in (100, 150):
        add(4960, L, 4950, 4960 + L + 10, 2000, F)
        add(10, L, -5, 10 + L + 3, 2000, F | wl.TF_CLAMP)
        add(len(genome) - L - 10, L, len(genome) - L - 14, len(genome) + 5, 2000, F | wl.TF_CLAMP)
        add(7000, L, 7004, 7000 + L - 5, 1000, F)
        add(7000, L, 7000, 7000 + L + 700, 3000, F)            # wide: unlimited by the Java rule
        add(7000, L, 6990, 7000 + L + 900, 3000, F | wl.TF_RAW_LIMITED)   # wider than 512 columns: generic kernel
    reads = np.concatenate(reads_l)
    tasks = np.array(tasks_l, dtype=wl.TASK_DTYPE)
    _compare(oracle, msa, reads, genome, tasks)
    _compare(oracle, msa, reads, genome, tasks, bw=20)


def test_long_reads_in_wide_windows(oracle, msa):
    """configs[4]'s pieces (`maxlen=500`) and the longest reads the aligner takes: 450-600 rows in windows of 460-700 columns (indels up to 60) — wider than the
    register-tiled kernels' 512 columns, so the limited fills run on the strip kernel up to 768 columns and on the row-sequential kernel beyond."""
    genome = wl.random_genome(80000, seed=61)
    for tight, ratio, n in ((True, 0.56, 500), (False, 0.336, 300)):
        reads, tasks = wl.make_msa_tasks(genome, n, seed=62 + tight, lengths=(450, 500, 600), flags=wl.TF_SCORE | wl.TF_TRACEBACK, tight=tight, ratio=ratio,
                                         max_indel=60, pad=8)
        cols = tasks["ref_end"] - tasks["ref_start"] + 1
        assert (cols > 512).mean() > 0.5 and cols.max() <= 768
        exp = _compare(oracle, msa, reads, genome, tasks)
        assert (exp["match_len"] > 0).sum() > 0.5 * n
    # a few windows beyond 768 columns as well (row-sequential class)
    reads, tasks = wl.make_msa_tasks(genome, 60, seed=64, lengths=(600,), flags=wl.TF_SCORE | wl.TF_TRACEBACK | wl.TF_RAW_LIMITED, tight=False, ratio=0.336, pad=130)
    assert ((tasks["ref_end"] - tasks["ref_start"] + 1) > 768).all()
    _compare(oracle, msa, reads, genome, tasks)


def test_narrow_candidates_of_the_generic_class(oracle, msa):
    """600-bp reads in 605-column windows are too wide for the tiled kernels (row-sequential class) yet narrow enough for the narrow kernel.
    The ones it finishes leave their reserved slot in the class list unused: the row-sequential kernel must take the list length from the
    device cursor, not from the classifier's count — checked with a small batch right after a large one (stale ids beyond the batch)."""
    genome = wl.random_genome(60000, seed=41).copy()
    big_reads, big_tasks = wl.make_msa_tasks(genome, 6000, seed=42, flags=wl.TF_SCORE | wl.TF_TRACEBACK)
    _compare(oracle, msa, big_reads, genome, big_tasks)
    rng = np.random.Generator(np.random.PCG64(43))
    reads_l, tasks_l, off = [], [], 0
    L = 600
    for k in range(12):
        pos = int(rng.integers(500, len(genome) - L - 500))
        r = genome[pos:pos + L].copy()
        if k % 3 == 1:
            r[300] = wl.ACGT[(int(np.searchsorted(wl.ACGT, r[300])) + 1) % 4]            # one substitution: still a narrow success
        if k % 3 == 2:
            r = np.concatenate([r[:200], genome[pos + 230:pos + 230 + 400]])                # 30-base deletion: the narrow kernel hands it over
        reads_l.append(r)
        tasks_l.append((off, 0, L, len(genome), pos - 2, pos + L - 1 + 3, wl.max_quality(L) - (400 if k % 3 != 2 else 3000), wl.TF_SCORE | wl.TF_TRACEBACK))
        off += L
    _compare(oracle, msa, np.concatenate(reads_l), genome, np.array(tasks_l, dtype=wl.TASK_DTYPE))


def test_reference_kind_agrees(oracle, msa):
    """Same comparison against the reference's own C (oracle/_ref), when it was built."""
    if not oracle.has_reference:
        pytest.skip("oracle/_ref not present")
    genome = wl.random_genome(50000, seed=24)
    reads, tasks = wl.make_msa_tasks(genome, 1500, seed=51, flags=wl.TF_SCORE | wl.TF_TRACEBACK)
    _compare(oracle, msa, reads, genome, tasks, kind="reference")


def test_invalid_and_empty(msa):
    genome = wl.random_genome(1000, seed=25)
    d_ref = msa.load_reference(genome)
    outs, _ = msa.align_batch(np.zeros(4, np.int8), d_ref, np.zeros(0, wl.TASK_DTYPE))
    assert len(outs) == 0
    tasks = np.zeros(2, wl.TASK_DTYPE)
    tasks[0] = (0, 0, 0, 1000, 10, 50, 100, 0)          # empty read
    tasks[1] = (0, 0, 4, 1000, 60, 50, 100, 0)          # refEnd < refStart
    outs, _ = msa.align_batch(np.frombuffer(b"ACGT", np.int8), d_ref, tasks)
    assert (outs["status"] == -3).all()
    msa.free(d_ref)


def test_score_no_indels(oracle, msa):
    """a10: ungapped site scoring, incl. N, sites hanging over both ends of the reference array, and match strings."""
    genome = wl.random_genome(30000, seed=26).copy()
    genome[100:130] = ord("N")
    reads, tasks = wl.make_msa_tasks(genome, 4000, seed=27, flags=0)
    nt = np.zeros(len(tasks) + 40, wl.NOINDEL_TASK_DTYPE)
    nt["read_off"][: len(tasks)] = tasks["read_off"]; nt["read_len"][: len(tasks)] = tasks["read_len"]
    nt["ref_len"] = len(genome); nt["ref_start"][: len(tasks)] = tasks["ref_start"] + 4
    for k in range(40):                      # overhanging / N-block sites
        j = len(tasks) + k
        nt["read_off"][j] = tasks["read_off"][k]; nt["read_len"][j] = tasks["read_len"][k]
        nt["ref_start"][j] = [-7, -1, len(genome) - 30, len(genome) - 1, 90, 120][k % 6]
    for flags in (0, 1):
        nt["flags"] = flags
        moff = np.zeros(len(nt) + 1, np.int64); np.cumsum(nt["read_len"], out=moff[1:])
        exp, em = oracle.noindel_batch(reads, genome, nt, match_off=moff if flags else None)
        d_ref = msa.load_reference(genome)
        got, gm = msa.scoreNoIndels(reads, d_ref, nt, match_off=moff if flags else None)
        msa.free(d_ref)
        assert np.array_equal(got, exp)
        if flags:
            assert gm.tobytes() == em.tobytes() and (exp == -99999).any()
    # the generator's vectorised scoreNoIndels agrees with the restatement on in-bounds sites
    L = 150
    sel = np.nonzero(tasks["read_len"] == L)[0][:500]
    r2 = np.stack([reads[tasks["read_off"][i]: tasks["read_off"][i] + L] for i in sel])
    g2 = np.stack([genome[tasks["ref_start"][i] + 4: tasks["ref_start"][i] + 4 + L] for i in sel])
    assert np.array_equal(wl.score_no_indels_batch(r2, g2), exp[sel])


def _spliced_case(rng, genome, L):
    """A read made of 2-3 exons separated by introns of 300-2500 bp, and the SiteScore-style gap array for it."""
    nex = int(rng.integers(2, 4))
    cuts = np.sort(rng.choice(np.arange(25, L - 25), size=nex - 1, replace=False))
    lens = np.diff(np.concatenate([[0], cuts, [L]]))
    pos = int(rng.integers(9000, len(genome) - 12000))
    exons, gaps, p = [], [], pos
    for i, ln in enumerate(lens):
        exons.append(genome[p:p + ln].copy()); gaps += [p, p + ln - 1]
        p += ln + int(rng.integers(300, 2500))
    read = np.concatenate(exons)
    for q in rng.integers(0, L, size=int(rng.integers(0, 3))):
        read[q] = wl.ACGT[rng.integers(0, 4)]
    return read, np.array(gaps, np.int32)


def test_gapped_reference_parity(oracle):
    """a15: makeGref + fill on the gapped reference + translateFromGappedCoordinate + traceback with '-' expansion, against the
    restatement (oracle/msa_oracle.c make_gref/from_gapped, MultiStateAligner11tsJNI.java:668-801), mixed with ungapped tasks."""
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    rng = np.random.Generator(np.random.PCG64(21))
    genome = wl.random_genome(60000, seed=8)
    reads, gt, gaps_all, cases = [], [], [], []
    off = 0
    for i in range(160):
        L = int(rng.choice([100, 150, 250]))
        if i % 4 == 3:      # ungapped task in the same batch
            p = int(rng.integers(9000, 40000)); read = genome[p:p + L].copy(); g = np.zeros(0, np.int32); lo, hi = p, p + L - 1
        else:
            read, g = _spliced_case(rng, genome, L); lo, hi = int(g[0]), int(g[-1])
        pad = int(rng.integers(0, 12))
        ms = int(0.3 * wl.max_quality(L)) if i % 5 else 0
        t = np.zeros(1, wl.GAPPED_TASK_DTYPE)
        t["t"]["read_off"] = off; t["t"]["ref_off"] = 0; t["t"]["read_len"] = L; t["t"]["ref_len"] = len(genome)
        t["t"]["ref_start"] = lo - pad; t["t"]["ref_end"] = hi + pad; t["t"]["min_score"] = ms
        t["t"]["flags"] = wl.TF_CLAMP | wl.TF_SCORE | wl.TF_TRACEBACK
        t["gaps_off"] = len(gaps_all); t["ngaps"] = len(g)
        gaps_all += g.tolist(); gt.append(t); reads.append(read); off += L
        cases.append((read, lo - pad, hi + pad, ms, g))
    gt = np.concatenate(gt); reads = np.concatenate(reads)
    cap = gt["t"]["read_len"].astype(np.int64) + 3002 + 128 * 30
    moff = np.zeros(len(gt) + 1, np.int64); np.cumsum(cap, out=moff[1:])
    msa = MultiStateAligner11tsCUDA(device=0)
    try:
        d_ref = msa.load_reference(genome)
        outs, mbuf = msa.align_batch_gapped(reads, d_ref, gt, np.array(gaps_all, np.int32), moff)
    finally:
        msa.close()
    ngapped_ok = 0
    for i, (read, a, b, ms, g) in enumerate(cases):
        sc, match, max4 = oracle.fill_and_score_limited_gapped(read, genome, a, b, ms, g)
        o = outs[i]
        assert o["status"] == 0, (i, o)
        if sc is None:
            assert o["score_len"] == 0 and o["result"][4] == 1, (i, o)
            continue
        assert o["score_len"] == len(sc) and o["score"][:len(sc)].tolist() == sc, (i, o["score"], sc)
        assert o["result"][:4].tolist() == max4.tolist()
        assert o["match_len"] == len(match) and mbuf[moff[i]:moff[i] + len(match)].tobytes() == match.tobytes(), i
        if len(g):
            ngapped_ok += 1
            assert (match == ord("D")).sum() >= 128          # an intron went through '-' symbols
    assert ngapped_ok >= 100


def test_large_batch_routings_agree(oracle):
    """200 k alignments of the bench workload (G4): the strip routing, the tiled routing and the mixed one are three different
    decompositions of the same fill — all results and match strings must be byte-identical; a 10 k sample is also checked against
    the oracle.  Covers multi-task-per-thread refill and the longest-first ordering, which small batches do not exercise."""
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    genome = wl.random_genome(300000, seed=77)
    reads, tasks = wl.make_msa_tasks(genome, 200000, seed=78, flags=wl.TF_SCORE | wl.TF_TRACEBACK)
    moff = wl.match_offsets(tasks)
    res = []
    for strip, narrow in ((16, 1000), (0, 1), (3, 1), (16, 0)):
        m = MultiStateAligner11tsCUDA()
        try:
            m.set_option("strip", strip); m.set_option("narrow", narrow)
            d_ref = m.load_reference(genome)
            outs, mb = m.align_batch(reads, d_ref, tasks, match_off=moff)
            res.append((outs.tobytes(), mb.tobytes(), m.stat("strip_tasks")))
        finally:
            m.close()
    assert res[0][2] > 50000 and res[1][2] == 0 and res[3][2] > 150000
    for r in res[1:]:
        assert r[0] == res[0][0] and r[1] == res[0][1]
    n = 10000
    exp, emb, _ = oracle.run_batch(reads, genome, tasks[:n], match_off=moff[:n + 1], threads=8)
    got = np.frombuffer(res[0][0], dtype=wl.OUT_DTYPE)[:n]
    assert got.tobytes() == exp.tobytes()
    gm = np.frombuffer(res[0][1], dtype=np.int8)
    for i in range(n):
        k = exp[i]["match_len"]
        if k > 0:
            assert gm[moff[i]:moff[i] + k].tobytes() == emb[moff[i]:moff[i] + k].tobytes()
