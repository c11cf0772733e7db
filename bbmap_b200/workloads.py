"""Synthetic inputs for the mapping-core kernels (generators are ours; RandomReads3 needs a JVM).

SURVEY.md §8(d) names the workloads.  G4 is the MultiStateAligner11ts microbenchmark of
BASELINE.json configs[2]: (read, candidate window) pairs with read length in {100,150,250},
70% true locus with ~1% substitutions, 20% true locus with one 1-40 bp indel, 10% unrelated
locus (exercises the fail path); window = locus +- SLOW_ALIGN_PADDING (4, reference
current/align2/BBMap.java:57).  minScore follows BBMapThread.scoreSlow
(current/align2/BBMapThread.java:255-260,306): max(scoreNoIndels, (int)(ratio*maxQ) - CLEARZONE1e).
All randomness is numpy PCG64 with the stated seed.
"""
import numpy as np

TASK_DTYPE = np.dtype([("read_off", "<i8"), ("ref_off", "<i8"), ("read_len", "<i4"), ("ref_len", "<i4"),
                       ("ref_start", "<i4"), ("ref_end", "<i4"), ("min_score", "<i4"), ("flags", "<i4")], align=True)
OUT_DTYPE = np.dtype([("result", "<i4", (5,)), ("path", "<i4"), ("iterations", "<i8"), ("score", "<i4", (8,)),
                      ("score_len", "<i4"), ("match_len", "<i4"), ("status", "<i4"), ("pad_", "<i4")], align=True)

TF_RAW_LIMITED, TF_RAW_UNLIMITED, TF_CLAMP, TF_SCORE, TF_TRACEBACK = 1, 2, 4, 8, 16
GAPPED_TASK_DTYPE = np.dtype([("t", TASK_DTYPE), ("gaps_off", "<i4"), ("ngaps", "<i4")], align=True)     # bbm_gapped_task, 48 bytes
NOINDEL_TASK_DTYPE = np.dtype([("read_off", "<i8"), ("ref_off", "<i8"), ("read_len", "<i4"), ("ref_len", "<i4"),
                               ("ref_start", "<i4"), ("flags", "<i4")], align=True)

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
CLEARZONE1e = 258          # current/align2/AbstractMapThread.java:142
SLOW_ALIGN_PADDING = 4     # current/align2/BBMap.java:57
POINTS_MATCH, POINTS_MATCH2 = 70, 100


def max_quality(n):
    """MultiStateAligner11tsJNI.maxQuality(int): MATCH + (n-1)*MATCH2."""
    return POINTS_MATCH + (n - 1) * POINTS_MATCH2


def random_genome(length, seed=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    return ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def score_no_indels_batch(reads2d, refs2d):
    """Vectorised MultiStateAligner11tsJNI.scoreNoIndels (…JNI.java:1033-1089) for in-bounds windows.
    reads2d, refs2d: (n, L) uint8.  Returns int32 scores."""
    n, L = reads2d.shape
    score = np.zeros(n, np.int32)
    mode = np.full(n, -1, np.int8)      # -1 none, 0 MS, 3 SUB
    tim = np.zeros(n, np.int32)
    N = ord("N")
    for i in range(L):
        c = reads2d[:, i]; r = refs2d[:, i]
        match = (c == r) & (c != N)
        nocall = ~match & (c == N)
        noref = ~match & ~nocall & (r == N)
        sub = ~match & ~nocall & ~noref
        # match
        cont = match & (mode == 0)
        score += np.where(cont, POINTS_MATCH2, 0).astype(np.int32)
        score += np.where(match & ~cont, POINTS_MATCH, 0).astype(np.int32)
        tim = np.where(cont, tim + 1, np.where(match, 0, tim))
        mode = np.where(match, 0, mode)
        # sub
        scont = sub & (mode == 3)
        tim = np.where(scont, tim + 1, np.where(sub, 0, tim))
        t1 = tim + 1
        pts = np.where(t1 > 5, -25, np.where(t1 > 1, -51, -127))
        score += np.where(sub, pts, 0).astype(np.int32)
        mode = np.where(sub, 3, mode)
    return score


def make_msa_tasks(genome, n, seed=2, block=100_000, **kw):
    """G4: returns (reads uint8[], tasks TASK_DTYPE[n]).  ref_off=0, ref_len=len(genome) for every task.  Generated in blocks of
    `block` tasks seeded (seed, block index), so the first k blocks of a large batch equal a smaller batch with the same seed."""
    if n <= block:
        return _make_msa_tasks_block(genome, n, seed, **kw)
    reads, tasks, off = [], [], 0
    for b in range((n + block - 1) // block):
        r, t = _make_msa_tasks_block(genome, min(block, n - b * block), seed if b == 0 else [seed, b], **kw)
        t["read_off"] += off; off += len(r)
        reads.append(r); tasks.append(t)
    return np.concatenate(reads), np.concatenate(tasks)


def _make_msa_tasks_block(genome, n, seed=2, lengths=(100, 150, 250), ratio=0.56, flags=TF_SCORE | TF_TRACEBACK,
                          frac_indel=0.2, frac_unrelated=0.1, sub_rate=0.01, n_rate=0.0005, pad=SLOW_ALIGN_PADDING,
                          max_indel=40, tight=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    G = len(genome)
    tasks = np.zeros(n, TASK_DTYPE)
    which = rng.integers(0, len(lengths), size=n)
    lens = np.asarray(lengths, np.int32)[which]
    read_off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=read_off[1:])
    reads = np.empty(int(read_off[-1]), np.uint8)
    kind = rng.random(n)
    is_unrel = kind < frac_unrelated
    is_indel = (~is_unrel) & (kind < frac_unrelated + frac_indel)
    margin = pad + max_indel + 64
    for L in lengths:
        idx = np.nonzero(lens == L)[0]
        m = len(idx)
        if m == 0:
            continue
        pos = rng.integers(margin, G - L - margin, size=m)
        ar = np.arange(L, dtype=np.int64)[None, :]
        # indels: +d = deletion from the read (ref span grows), -d = insertion into the read (ref span shrinks)
        ind = is_indel[idx]
        dlen = np.where(ind, rng.integers(1, max_indel + 1, size=m), 0)
        is_del = rng.random(m) < 0.5
        kpos = rng.integers(10, L - 10 - 1, size=m)[:, None]
        d = dlen[:, None]
        ilen = np.minimum(d, L - 20 - 1)  # keep insertions inside the read
        src_del = ar + np.where(ar >= kpos, d, 0)
        src_ins = ar - np.clip(ar - kpos, 0, ilen)
        src = np.where(is_del[:, None], src_del, src_ins)
        base = genome[pos[:, None] + src]
        inserted = (~is_del[:, None]) & (ar >= kpos) & (ar < kpos + ilen) & ind[:, None]
        rnd = ACGT[rng.integers(0, 4, size=(m, L), dtype=np.uint8)]
        base = np.where(inserted, rnd, base)
        # substitutions
        submask = rng.random((m, L)) < sub_rate
        code = np.searchsorted(ACGT, base)  # ACGT is sorted: A<C<G<T
        sub = ACGT[(code + rng.integers(1, 4, size=(m, L))) % 4]
        base = np.where(submask, sub, base)
        # unrelated reads: fully random
        un = is_unrel[idx]
        base = np.where(un[:, None], rnd, base)
        # a few no-calls
        base = np.where(rng.random((m, L)) < n_rate, np.uint8(ord("N")), base)
        span = np.where(ind, np.where(is_del, L + dlen, L - np.minimum(dlen, L - 21)), L)
        a = pos - pad
        b = pos + span - 1 + pad
        maxq = max_quality(L)
        min_limit = int(ratio * maxq) - CLEARZONE1e
        if tight:
            ni = score_no_indels_batch(base, genome[pos[:, None] + ar])
            ms = np.maximum(ni, min_limit)
        else:
            ms = np.full(m, min_limit, np.int32)
        flat = (read_off[idx][:, None] + ar).ravel()
        reads[flat] = base.ravel()
        tasks["read_len"][idx] = L
        tasks["ref_start"][idx] = a
        tasks["ref_end"][idx] = b
        tasks["min_score"][idx] = ms
    tasks["read_off"] = read_off[:-1]
    tasks["ref_off"] = 0
    tasks["ref_len"] = G
    tasks["flags"] = flags
    return reads, tasks


def match_offsets(tasks, extra=0):
    """Match-string slot per task: rows + columns (+extra) bytes (Java allocates rows+cols-1, …JNI.java:380)."""
    a = tasks["ref_start"].astype(np.int64); b = tasks["ref_end"].astype(np.int64)
    clamp = (tasks["flags"] & TF_CLAMP) != 0
    a = np.where(clamp, np.maximum(a, 0), a)
    b = np.where(clamp, np.minimum(b, tasks["ref_len"].astype(np.int64) - 1), b)
    cap = tasks["read_len"].astype(np.int64) + np.maximum(b - a + 1, 0) + extra
    cap = (cap + 3) & ~np.int64(3)
    off = np.zeros(len(tasks) + 1, np.int64)
    np.cumsum(cap, out=off[1:])
    return off


# ----------------------------------------------------------------------------------------------------------------
# G6: BandedAligner tasks: (query, ref) pairs as Dedupe would pass them (reads / contigs), 0-3% edits,
# maxEdits in {2,5,26}, maxWidth = max(min(bw, 2*maxEdits+1), 3) | 1 with Dedupe's default bw=9 (jgi/Dedupe.java:5715-5717)
# unless `widths` overrides it; all four directions; exact in {0,1}.
BAND_TASK_DTYPE = np.dtype([("query_off", "<i8"), ("ref_off", "<i8"), ("query_len", "<i4"), ("ref_len", "<i4"), ("qstart", "<i4"),
                            ("rstart", "<i4"), ("max_edits", "<i4"), ("max_width", "<i4"), ("exact", "<i4"), ("dir", "<i4")], align=True)
BAND_OUT_DTYPE = np.dtype([("edits", "<i4"), ("rv", "<i4", (5,)), ("status", "<i4"), ("pad_", "<i4")], align=True)
DIR_FORWARD, DIR_FORWARD_RC, DIR_REVERSE, DIR_REVERSE_RC = 0, 1, 2, 3
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def revcomp(a):
    return _COMP[a[::-1]]


def make_banded_tasks(n, seed=6, min_len=150, max_len=5000, edit_rate=0.03, max_edits=(2, 5, 26), bw=9, widths=None,
                      n_rate=0.001, ragged=True):
    """Returns (queries uint8[], refs uint8[], tasks BAND_TASK_DTYPE[n])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tasks = np.zeros(n, BAND_TASK_DTYPE)
    qs, rs = [], []
    qoff = roff = 0
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        ref = ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        q = ref.copy()
        rate = float(rng.random()) * edit_rate
        ne = int(rng.binomial(L, rate))
        for _ in range(ne):
            p = int(rng.integers(0, len(q)))
            k = rng.random()
            if k < 0.6:
                q[p] = ACGT[(np.searchsorted(ACGT, q[p]) + rng.integers(1, 4)) % 4]
            elif k < 0.8 and len(q) > 20:
                q = np.delete(q, p)
            else:
                q = np.insert(q, p, ACGT[rng.integers(0, 4)])
        if n_rate > 0:
            m = rng.random(len(q)) < n_rate
            q = np.where(m, np.uint8(ord("N")), q)
        if ragged and rng.random() < 0.3:          # unequal lengths: exercises the query/ref swap rules
            cut = int(rng.integers(1, max(2, L // 4)))
            if rng.random() < 0.5:
                q = q[:-cut] if len(q) > cut + 10 else q
            else:
                ref = ref[:-cut] if len(ref) > cut + 10 else ref
        d = int(rng.integers(0, 4))
        me = int(max_edits[int(rng.integers(0, len(max_edits)))])
        if widths is None:
            mw = max(min(bw, 2 * me + 1), 3) | 1
        else:
            mw = int(widths[int(rng.integers(0, len(widths)))])
        if d in (DIR_FORWARD_RC, DIR_REVERSE_RC):
            q = revcomp(q)
        ql, rl = len(q), len(ref)
        if d == DIR_FORWARD:
            qstart, rstart = 0, 0
        elif d == DIR_FORWARD_RC:
            qstart, rstart = ql - 1, 0
        elif d == DIR_REVERSE:
            qstart, rstart = ql - 1, rl - 1
        else:
            qstart, rstart = 0, rl - 1
        if rng.random() < 0.15:                     # interior starts
            sh = int(rng.integers(0, 20))
            if d == DIR_FORWARD:
                qstart, rstart = sh, sh
            elif d == DIR_REVERSE:
                qstart, rstart = ql - 1 - sh, rl - 1 - sh
        tasks[i] = (qoff, roff, ql, rl, qstart, rstart, me, mw, int(rng.random() < 0.5), d)
        qs.append(q); rs.append(ref)
        qoff += ql; roff += rl
    return np.concatenate(qs), np.concatenate(rs), tasks


# ----------------------------------------------------------------------------------------------------------------
# Read batches for the seeding kernels: bases + phred qualities (offset removed), as Read.validate leaves them
# (stream/Read.java:81-215: qualities clamped to [2,41] for ACGT, 0 for undefined bases).
def make_read_batch(n, seed=9, lengths=(100, 150, 250), flat_q=None, n_rate=0.002, lowq_tail=0.3):
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.asarray(lengths, np.int64)[rng.integers(0, len(lengths), size=n)]
    off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    bases = ACGT[rng.integers(0, 4, size=total, dtype=np.uint8)]
    if flat_q is not None:
        qual = np.full(total, flat_q, np.uint8)
    else:
        qual = rng.integers(2, 42, size=total).astype(np.uint8)
        hi = rng.random(total) < 0.8
        qual = np.where(hi, np.maximum(qual, 30), qual).astype(np.uint8)
        # degrade read tails like Illumina reads
        for r in np.nonzero(rng.random(n) < lowq_tail)[0]:
            L = int(lens[r]); t = int(rng.integers(5, max(6, L // 2)))
            qual[off[r] + L - t: off[r] + L] = rng.integers(2, 12, size=t)
    nmask = rng.random(total) < n_rate
    bases = np.where(nmask, np.uint8(ord("N")), bases)
    qual = np.where(nmask, 0, qual).astype(np.uint8)
    # a few reads that are mostly N / shorter than k
    return bases, qual, off


# ----------------------------------------------------------------------------------------------------------------
# G2 / G3: paired 2x150 reads drawn from a packed reference — insert size uniform 200-500, each base substituted
# with probability `sub_rate`, and with probability `indel_rate*L` one 1-3 bp insertion or deletion per read; Q=30 flat
# (Shared.FAKE_QUAL).  Mate 2 is the reverse complement of the far end of the fragment (Illumina FR).  Vectorised numpy.
def make_mapping_reads(chrom_bytes, chrom_off, table, npairs, L=150, seed=2, sub_rate=0.01, indel_rate=0.01 / 3, insert=(200, 500), qual=30):
    """chrom_bytes/chrom_off/table as returned by bbmap_b200.index.pack_chromosomes.  Returns dict(bases, qual, off, truth) where
    reads 2i and 2i+1 are the mates of pair i and truth[r] = (chrom, strand, start, stop) in chromosome coordinates."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tab = np.asarray(table, np.int64)                       # (chrom 1-based, start, length)
    w = tab[:, 2].astype(np.float64); w /= w.sum()
    sc = rng.choice(len(tab), size=npairs, p=w)
    ins = rng.integers(insert[0], insert[1] + 1, size=npairs)
    ins = np.minimum(ins, tab[sc, 2] - 8)
    fstart = tab[sc, 1] + (rng.random(npairs) * (tab[sc, 2] - ins - 4)).astype(np.int64)     # fragment start in the chromosome
    fstrand = rng.integers(0, 2, size=npairs)
    n = 2 * npairs
    # genomic start of each read's footprint on the plus strand
    start = np.empty(n, np.int64); strand = np.empty(n, np.int64); chrom = np.repeat(tab[sc, 0], 2)
    left, right = fstart, fstart + ins - L
    start[0::2] = np.where(fstrand == 0, left, right); strand[0::2] = fstrand
    start[1::2] = np.where(fstrand == 0, right, left); strand[1::2] = 1 - fstrand
    return _reads_from_footprints(chrom_bytes, chrom, start, strand, L, rng, sub_rate, indel_rate, qual, chrom_off)


def _reads_from_footprints(chrom_bytes, chrom, start, strand, L, rng, sub_rate, indel_rate, qual, chrom_off=None):
    n = len(start)
    j = np.arange(L, dtype=np.int64)[None, :]
    has_indel = rng.random(n) < indel_rate * L
    is_del = rng.random(n) < 0.5
    d = rng.integers(1, 4, size=n)
    q = rng.integers(20, L - 20, size=n)
    shift = np.where((has_indel & is_del)[:, None] & (j >= q[:, None]), d[:, None], 0)
    shift = shift - np.where((has_indel & ~is_del)[:, None] & (j >= (q + d)[:, None]), d[:, None], 0)
    idx = start[:, None] + j + shift
    base0 = np.zeros(n, np.int64) if chrom_off is None else np.asarray(chrom_off, np.int64)[chrom - 1]
    reads = chrom_bytes[np.minimum(idx + base0[:, None], len(chrom_bytes) - 1)]
    insmask = (has_indel & ~is_del)[:, None] & (j >= q[:, None]) & (j < (q + d)[:, None])
    reads = np.where(insmask, ACGT[rng.integers(0, 4, size=(n, L), dtype=np.uint8)], reads)
    sub = rng.random((n, L)) < sub_rate
    reads = np.where(sub, ACGT[(np.searchsorted(ACGT, np.minimum(reads, ord("T"))) + rng.integers(1, 4, size=(n, L))) % 4], reads).astype(np.uint8)
    span = L + np.where(has_indel, np.where(is_del, d, -d), 0)
    minus = strand == 1
    reads[minus] = revcomp(reads[minus].reshape(-1)).reshape(-1, L)[::-1]
    truth = np.stack([chrom, strand, start, start + span - 1], axis=1).astype(np.int32)
    off = np.arange(n + 1, dtype=np.int64) * L
    return {"bases": reads.reshape(-1), "qual": np.full(n * L, qual, np.uint8), "off": off, "truth": truth}


def make_spliced_reads(chrom_bytes, chrom_off, table, npairs, L=150, seed=7, intron=(300, 12000), frac=0.5, sub_rate=0.005, insert=(200, 500), qual=30):
    """RNA-seq-style pairs (BASELINE configs[4] shape within the default maxindel of 16000): in a fraction `frac` of the pairs the left read of the
    fragment is spliced over one intron of `intron` bp (its footprint on the reference is L + intron bases, the mate lies behind it).  Returns
    dict(bases, qual, off, truth, spliced): reads 2i / 2i+1 are mates; truth[r] = (chrom, strand, start, stop) in chromosome coordinates."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tab = np.asarray(table, np.int64)
    w = tab[:, 2].astype(np.float64); w /= w.sum()
    sc = rng.choice(len(tab), size=npairs, p=w)
    ins = rng.integers(insert[0], insert[1] + 1, size=npairs)
    gap = np.where(rng.random(npairs) < frac, rng.integers(intron[0], intron[1] + 1, size=npairs), 0)
    span = np.minimum(ins + gap, tab[sc, 2] - 8)
    gap = np.where(span < ins + gap, 0, gap); span = ins + gap
    fstart = tab[sc, 1] + (rng.random(npairs) * (tab[sc, 2] - span - 4)).astype(np.int64)
    fstrand = rng.integers(0, 2, size=npairs)
    q = rng.integers(40, L - 40, size=npairs)
    j = np.arange(L, dtype=np.int64)[None, :]
    base0 = np.asarray(chrom_off, np.int64)[tab[sc, 0] - 1]
    left_idx = fstart[:, None] + j + np.where(j >= q[:, None], gap[:, None], 0)
    right_start = fstart + span - L
    right_idx = right_start[:, None] + j
    left = chrom_bytes[np.minimum(left_idx + base0[:, None], len(chrom_bytes) - 1)]
    right = chrom_bytes[np.minimum(right_idx + base0[:, None], len(chrom_bytes) - 1)]
    for arr in (left, right):
        sub = rng.random(arr.shape) < sub_rate
        arr[sub] = ACGT[(np.searchsorted(ACGT, np.minimum(arr[sub], ord("T"))) + rng.integers(1, 4, size=int(sub.sum()))) % 4]
    right = revcomp(right.reshape(-1)).reshape(-1, L)[::-1]            # the right read of a fragment is its minus-strand end
    n = 2 * npairs
    reads = np.empty((n, L), np.uint8); truth = np.empty((n, 4), np.int32)
    plus = fstrand == 0
    # fragment on the plus strand: read 1 = left (+), read 2 = right (-); on the minus strand the roles swap
    reads[0::2] = np.where(plus[:, None], left, right); reads[1::2] = np.where(plus[:, None], right, left)
    lt = np.stack([tab[sc, 0], np.zeros(npairs, np.int64), fstart, fstart + L + gap - 1], axis=1)
    rt = np.stack([tab[sc, 0], np.ones(npairs, np.int64), right_start, right_start + L - 1], axis=1)
    truth[0::2] = np.where(plus[:, None], lt, rt); truth[1::2] = np.where(plus[:, None], rt, lt)
    spl = np.zeros(n, bool); spl[0::2] = plus & (gap > 0); spl[1::2] = ~plus & (gap > 0)
    off = np.arange(n + 1, dtype=np.int64) * L
    return {"bases": reads.reshape(-1), "qual": np.full(n * L, qual, np.uint8), "off": off, "truth": truth, "spliced": spl}


def make_long_reads(chrom_bytes, chrom_off, table, nreads, L=1000, seed=8, sub_rate=0.01, indel_rate=0.0005, qual=30):
    """Single-ended long reads (BASELINE configs[4]: "1 kbp long reads"): L bases from a random locus and strand, ~1 % substitutions, a 1-3 bp indel in
    about half of them.  The mapper takes them after bbmap_b200.reads.break_reads(..., max_len=500) (`maxlen=500`, ReformatReads.breakReads).
    Returns dict(bases, qual, off, truth, names, name_off)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tab = np.asarray(table, np.int64)
    w = tab[:, 2].astype(np.float64); w /= w.sum()
    sc = rng.choice(len(tab), size=nreads, p=w)
    start = tab[sc, 1] + (rng.random(nreads) * np.maximum(tab[sc, 2] - L - 8, 1)).astype(np.int64)
    strand = rng.integers(0, 2, size=nreads)
    R = _reads_from_footprints(chrom_bytes, tab[sc, 0], start, strand, L, rng, sub_rate, indel_rate, qual, chrom_off)
    names = [b"long_%d" % i for i in range(nreads)]
    R["names"] = np.frombuffer(b"".join(names), np.int8).copy()
    R["name_off"] = np.zeros(nreads + 1, np.int64); np.cumsum([len(x) for x in names], out=R["name_off"][1:])
    return R
