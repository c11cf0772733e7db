"""TEST INFRASTRUCTURE — a second, independent restatement of the reference's index construction and analysis, in numpy, written
from the Java text alone; shares no code with oracle/index_oracle.c (the C restatement the CUDA index build is tested against).

Follows:
  IndexMaker4.CountThread.countSizes / fillArrays   current/align2/IndexMaker4.java:204-421  (k-mers of one block, period-<=2 ban, order of a list)
  BBIndex.toNumber / setChromBits                   current/align2/BBIndex.java:3038-3057,3148-3164
  BBIndex.analyzeIndex                              current/align2/BBIndex.java:101-191      (COUNTS, clumpy keys, lengthHistogram, derived limits)
  Tools.makeLengthHistogram3/4                      current/align2/Tools.java:1797-1850
  AminoAcid.reverseComplementBinaryFast             current/dna/AminoAcid.java:258-271
"""
import math

import numpy as np

CLUMPY_MAX_DIST = 5
CLUMPY_MIN_LENGTH_INDEX = 2000
CLUMPY_FRACTION = np.float32(0.75)
SMALL_GENOME_LIST = 20
DOUBLE_SEARCH_THRESH_MULT = np.float32(0.25)
BASE_POINTS_PER_SITE = -50


def _codes(chrom):
    lut = np.full(256, -1, np.int64)
    for i, c in enumerate("ACGT"):
        lut[ord(c)] = i
        lut[ord(c.lower())] = i
    lut[ord("U")] = 3
    lut[ord("u")] = 3
    return lut[np.asarray(chrom).view(np.uint8)]


def rcomp_keys(keys, k):
    keys = np.asarray(keys, np.int64)
    out = np.zeros_like(keys)
    x = keys.copy()
    for _ in range(k):
        out = (out << 2) | (3 - (x & 3))
        x >>= 2
    return out


def block_kmers(chroms, first_chrom, k, chrombits):
    """chroms: list of byte arrays (chromosome first_chrom, first_chrom+1, ... of one block).  Returns (keys, site numbers) in the order the
    reference fills them: chromosome by chromosome, positions ascending."""
    shift = 31 - chrombits
    low = (1 << chrombits) - 1
    allk, alls = [], []
    for ci, arr in enumerate(chroms):
        chrom = first_chrom + ci
        code = _codes(arr)
        n = len(code)
        max_index = n - 1
        stop = max_index - k + 1                 # the loop runs a < max: the last k-mer of a chromosome array is never indexed
        if stop <= 0:
            continue
        bad = (code < 0).astype(np.int64)
        cb = np.concatenate(([0], np.cumsum(bad)))
        a = np.arange(0, stop)
        defined = (cb[a + k] - cb[a]) == 0
        key = np.zeros(stop, np.int64)
        c0 = np.where(code < 0, 0, code)
        for j in range(k):
            key = (key << 2) | c0[a + j]
        banned = (key >> 4) == (key & ((1 << (2 * k - 4)) - 1))
        raw = np.asarray(arr).view(np.uint8)[a]
        owned = (raw == ord("A")) | (raw == ord("C")) | (raw == ord("G")) | (raw == ord("T"))   # `array[a]==idb`: each of the four threads takes the k-mers
        keep = defined & ~banned & owned                                                        # that START with its upper-case base; none takes a c g t u U
        allk.append(key[keep])
        alls.append((((chrom & low) << shift) | a[keep]).astype(np.int64))
    if not allk:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(allk), np.concatenate(alls)


def build_block(chroms, first_chrom, k, chrombits):
    keys, sites = block_kmers(chroms, first_chrom, k, chrombits)
    order = np.argsort(keys, kind="stable")
    starts = np.zeros((1 << (2 * k)) + 1, np.int64)
    np.cumsum(np.bincount(keys, minlength=1 << (2 * k)), out=starts[1:])
    return starts.astype(np.int32), sites[order].astype(np.int32)


def _wrap32(x):
    return ((int(x) + (1 << 31)) % (1 << 32)) - (1 << 31)


def length_histogram3(x, buckets=1000):
    mx = int(x.max())
    assert mx <= len(x)
    counts = np.bincount(x[x >= 0], minlength=mx + 1)
    total = int(x[x >= 0].astype(np.int64).sum())
    if total <= 0:
        total = sum(i * int(counts[i]) for i in range(1, len(counts)))
    hist = [0] * (buckets + 1)
    s = 0
    ptr = 0
    for i in range(buckets):
        nxt = ((total * i) + buckets // 2) // buckets
        while ptr < len(counts) and s < nxt:
            s += _wrap32(int(counts[ptr]) * ptr)         # int*int in Java
            ptr += 1
        hist[i] = max(0, ptr - 1)
    hist[buckets] = len(counts) - 1
    return hist


def analyze(blocks, k, fraction_to_exclude, max_average_list_to_search):
    """blocks: [(starts, sites)].  Returns (COUNTS, hist, MAX_USABLE_LENGTH, MAX_USABLE_LENGTH2, POINTS_PER_SITE)."""
    ks = 1 << (2 * k)
    counts = np.zeros(ks, np.int64)
    clumps = np.zeros(ks, np.int64)
    allkeys = np.arange(ks, dtype=np.int64)
    rc = rcomp_keys(allkeys, k)
    canon = np.minimum(allkeys, rc)
    for starts, sites in blocks:
        st = starts.astype(np.int64)
        counts = np.minimum(2 ** 31 - 1, counts + np.diff(st))
        if len(sites) > 1:
            dif = np.diff(sites.astype(np.int64))
            clumpy = (dif > 0) & (dif <= CLUMPY_MAX_DIST)
            # pair (i-1, i) belongs to a key only when both are inside its list: i is not the first element of a list
            first = np.zeros(len(sites), bool)
            first[st[:-1][st[:-1] < len(sites)]] = True
            clumpy &= ~first[1:]
            owner = np.searchsorted(st, np.arange(1, len(sites)), side="right") - 1
            np.add.at(clumps, canon[owner[clumpy]], 1)
    sym = counts.copy()
    lower = allkeys < rc
    tot = np.minimum(2 ** 31 - 1, counts[lower] + counts[rc[lower]])
    sym[allkeys[lower]] = tot
    sym[rc[lower]] = tot
    for key in np.nonzero(clumps > 0)[0]:
        ln = int(sym[key])
        if ln > CLUMPY_MIN_LENGTH_INDEX and np.float32(int(clumps[key])) > CLUMPY_FRACTION * np.float32(ln):
            sym[key] = 0
            sym[rc[key]] = 0
    hist = length_histogram3(sym.astype(np.int64))
    f = np.float32(fraction_to_exclude)
    i1 = int((np.float32(1) - f) * np.float32(len(hist) - 1))
    i2 = int((np.float32(1) - f * DOUBLE_SEARCH_THRESH_MULT) * np.float32(len(hist) - 1))
    mul = max(2 * SMALL_GENOME_LIST, hist[i1])
    mul2 = max(6 * SMALL_GENOME_LIST, hist[i2])
    pps = int(math.floor(float((np.float32(BASE_POINTS_PER_SITE) * np.float32(4000)) / np.float32(max(2 * SMALL_GENOME_LIST, hist[max_average_list_to_search])))))
    if pps == 0:
        pps = -1
    return sym.astype(np.int32), hist, mul, mul2, pps
