"""TEST INFRASTRUCTURE — a second, independent restatement of the reference's ingest and seeding rows, written from the Java
text alone in plain Python with numpy float32 scalars for Java `float` (one rounding per operation, as the JVM does).  It shares no code with
oracle/host_oracle.c (the C restatement the CUDA kernels are tested against); tests/test_seed_independent.py asserts both give the same answer.

Follows:
  Read.validate                       current/stream/Read.java:81-215 (switches :3406-3418)
  AminoAcid tables / reverse complement  current/dna/AminoAcid.java:105-133,203-211,258-271,567-647
  QualityTools tables                 current/align2/QualityTools.java:475-539
  QualityTools.makeKeyProbs           current/align2/QualityTools.java:188-280
  QualityTools.makeKeyScores / makeByteScoreArray   :125-180
  KeyRing.makeKeys / reverseComplementKeys / reverseOffsets / desiredKeysFromDensity / makeOffsets3   current/align2/KeyRing.java:23-45,125-137,269-282,396-506
  ChromosomeArray.toNumber            current/dna/ChromosomeArray.java:297-307
  AbstractMapThread.quickMap          current/align2/AbstractMapThread.java:643-733
  Read.avgQualityByProbability / expectedErrors / countUndefined   current/stream/Read.java:1738-1745,1992-1999,2115-2132
"""
import math

import numpy as np

F = np.float32

# ---------------- AminoAcid static tables ----------------
NUMBER_TO_BASE_EXT = " ACMGRSVTWYHKDBNX"
NUMBER_TO_COMP_EXT = " TGKCYWBASRDMHVNX"

BASE_TO_NUMBER = [-1] * 128
for _i, _c in enumerate("ACGT"):
    BASE_TO_NUMBER[ord(_c)] = _i
    BASE_TO_NUMBER[ord(_c.lower())] = _i
BASE_TO_NUMBER[ord("U")] = 3
BASE_TO_NUMBER[ord("u")] = 3

BASE_TO_NUMBER_EXT = [-1] * 128
for _i, _c in enumerate(NUMBER_TO_BASE_EXT):
    if not _c.isspace():
        BASE_TO_NUMBER_EXT[ord(_c)] = _i
        BASE_TO_NUMBER_EXT[ord(_c.lower())] = _i
BASE_TO_NUMBER_EXT[ord("U")] = 8
BASE_TO_NUMBER_EXT[ord("u")] = 8

BASE_TO_COMP_EXT = [-1] * 128
for _c, _d in zip(NUMBER_TO_BASE_EXT, NUMBER_TO_COMP_EXT):
    BASE_TO_COMP_EXT[ord(_c)] = ord(_d)
    BASE_TO_COMP_EXT[ord(_c.lower())] = ord(_d.lower())
BASE_TO_COMP_EXT[ord("U")] = ord("A")
BASE_TO_COMP_EXT[ord("u")] = ord("a")
for _c in "? -*.":
    BASE_TO_COMP_EXT[ord(_c)] = ord(_c)


def is_fully_defined(b):
    return 0 <= b < 128 and BASE_TO_NUMBER[b] >= 0


def reverse_complement_bases(bases):
    n = len(bases)
    return [BASE_TO_COMP_EXT[bases[n - 1 - i]] for i in range(n)]


# ---------------- Read.validate (nucleotide reads) ----------------
def validate(bases, quality, fix_junk=False, u_to_t=False, to_upper=False, lower_to_n=False, change_quality=True):
    """bases / quality: lists of ints (signed byte values), modified copies are returned with the junk flag."""
    b = list(bases)
    q = None if quality is None else list(quality)
    junk = False
    if u_to_t:
        for i in range(len(b)):
            if chr(b[i] & 0xFF).upper() == "U" and 0 <= b[i] < 128:
                b[i] = ord("t") if b[i] == ord("u") else ord("T")
    for i in range(len(b)):
        num = BASE_TO_NUMBER_EXT[b[i]] if 0 <= b[i] < 128 else -1
        if num < 0:
            if fix_junk:
                b[i] = ord("N")
            else:
                junk = True
                break
    NOCALL = ord("N")
    others = (ord("-"), ord("."), ord("X"))
    if q is not None:
        for i in range(len(b)):
            x = b[i]
            if change_quality:
                if is_fully_defined(x):
                    if q[i] < 2:
                        q[i] = 2
                    elif q[i] > 41:
                        q[i] = 41
                else:
                    q[i] = 0
                    if x in others or x == ord("n"):
                        b[i] = NOCALL
            elif not is_fully_defined(x):
                if x in others or x == ord("n"):
                    b[i] = NOCALL
            if to_upper and x > 90:
                b[i] -= 32
            elif lower_to_n and x > 90:
                b[i] = NOCALL
    elif to_upper:
        for i in range(len(b)):
            x = b[i]
            if x > 90:
                b[i] -= 32
            if x in others:
                b[i] = NOCALL
    elif lower_to_n:
        for i in range(len(b)):
            x = b[i]
            if x > 90 or x in others:
                b[i] = NOCALL
    else:
        for i in range(len(b)):
            if b[i] in others:
                b[i] = NOCALL
    return b, q, junk


# ---------------- QualityTools tables ----------------
def _tables():
    pe = [F(math.pow(10, 0 - .1 * i)) for i in range(127)]
    pe[0] = F(.8)
    pc = [F(1) - x for x in pe]
    with np.errstate(divide="ignore"):
        pci = [F(1) / x for x in pc]
    return pe, pc, pci


PROB_ERROR, PROB_CORRECT, PROB_CORRECT_INVERSE = _tables()


def java_round(x):
    """Math.round(float|double): floor(x + 1/2) evaluated exactly."""
    return int(math.floor(float(x) + 0.5))


def make_key_probs(quality, bases, k):
    """Probability that the k-mer starting at each position contains an error (usemodulo off)."""
    n = len(bases) - k + 1
    if quality is None:
        return [F(0)] * n
    out = [F(0)] * n
    key1 = F(1)
    since = 0
    for i in range(k):
        qv = quality[i]
        since = since + 1 if qv > 0 else 0
        key1 = key1 * PROB_CORRECT[qv]
    out[0] = F(1) - key1
    if since < k:
        out[0] = F(1)
    a = 0
    for bpos in range(k, len(quality)):
        qa, qb = quality[a], quality[bpos]
        since = since + 1 if qb > 0 else 0
        key1 = (key1 * PROB_CORRECT_INVERSE[qa]) * PROB_CORRECT[qb]
        out[a + 1] = F(1) - key1
        if since < k:
            out[a + 1] = F(1)
        a += 1
    return out


def desired_keys_from_density(readlen, k, density, min_keys):
    slots = readlen - k + 1
    desired = int(math.ceil(float((F(readlen) * F(density)) / F(k))))
    return min(slots, max(min_keys, desired))


def make_offsets3(prob, readlen_original, k, density, max_density, min_keys, semiperfect=False):
    readlen = readlen_original
    max_index = readlen - k
    left, right = 0, max_index
    lim2 = F(0.9999)
    lim1 = F(0.99) if semiperfect else F(0.94)
    while left <= right and prob[left] >= lim1:
        left += 1
    while right >= left and prob[right] >= lim1:
        right -= 1
    potential = sum(1 for i in range(left, right + 1) if prob[i] < lim2)
    if potential == 0 or right < left:
        return None
    readlen = right - left + k
    desired = desired_keys_from_density(readlen_original, k, density, min_keys)
    if readlen < readlen_original:
        desired = min(desired, desired_keys_from_density(readlen, k, max_density, min_keys))
    desired = min(desired, potential)
    offsets = []
    interval = F(right - left) / F(max(desired - 1, 1))
    interval_int = int(interval) + 1
    f = F(left)
    prev = -1
    j = left
    for _ in range(desired):
        x = -1
        if prev < j:
            if prob[j] < lim2 and (prev < 0 or j - prev > 0):
                x = j
            else:
                kk = j - 1
                while kk > prev + 2:
                    if prob[kk] < lim2:
                        x = kk
                        break
                    kk -= 1
                if x < 0:
                    kk = j + 1
                    stop = min(j + interval_int, right)
                    while kk < stop:
                        if prob[kk] < lim2:
                            x = kk
                            break
                        kk += 1
        if x > -1:
            offsets.append(x)
            prev = x
        else:
            prev = max(prev, j - 2)
        f = f + interval
        j = min(max_index, max(j + 1, java_round(f)))
    return offsets


def to_number(bases, a, b):
    out = 0
    for i in range(a, b + 1):
        x = BASE_TO_NUMBER[bases[i]] if 0 <= bases[i] < 128 else -1
        if x < 0:
            return -1
        out = (out << 2) | x
    return out


def rcomp_binary(kmer, k):
    """AminoAcid.reverseComplementBinaryFast, on the 2-bit symbols directly (the byte table of the reference is the same map four symbols at a
    time).  kmer == -1 (an undefined key) behaves as 32 one-bits under Java's arithmetic shift: every symbol complements to 0."""
    out = 0
    for _ in range(k):
        out = (out << 2) | ((~kmer) & 3)
        kmer >>= 2
    return out


def expected_errors(bases, quality):
    s = F(0)
    for b, q in zip(bases, quality):
        if is_fully_defined(b):
            s = s + PROB_ERROR[q]
    return s


def avg_quality_by_probability(bases, quality):
    if quality is None:
        return 40
    if len(quality) == 0:
        return 0
    p = expected_errors(bases, quality) / F(len(quality))
    prob = 1.0 - float(F(1) - p)            # probCorrectToPhred(1-p) -> probErrorToPhred(1-prob), the second subtraction in double
    if prob >= 1:
        phred = 0.0
    elif prob <= 0.000001:
        phred = 60.0
    else:
        phred = -10 * math.log10(prob)
    return min(41, max(0, java_round(phred)))


def quick_map_seed(bases, quality, k=13, max_desired_keys=15, base_key_hit_score=1300, min_hits_to_keep=1, key_density=1.9, max_key_density=3.0,
                   min_key_density=1.5):
    """The seeding half of quickMap.  Returns None when the read gets no seeds (quickMap returns 0 / -1), else
    dict(offsets, keys, keyScores, baseScores, offsetsM, keysM)."""
    L = len(bases)
    if L < k:
        return None
    undefined = sum(1 for b in bases if not (0 <= b < 128 and BASE_TO_NUMBER[b] >= 0))
    if undefined > 25 and L - undefined < undefined:
        return None
    den2 = F(max_desired_keys * k) / F(L)
    den2 = max(F(min_key_density), den2)
    den2 = min(F(key_density), den2, F(k))
    if L <= 50:
        den3 = F(max_key_density)
    elif L >= 200:
        den3 = F(max_key_density) - F(0.5)
    else:
        den3 = F(max_key_density) - F(0.003333333333) * F(L - 50)
    den3 = max(F(key_density), den3)
    den3 = min(F(k), den3)
    prob = make_key_probs(quality, bases, k)
    offsets = make_offsets3(prob, L, k, den2, den3, 2)
    if offsets is None or len(offsets) < min_hits_to_keep or (quality is not None and avg_quality_by_probability(bases, quality) < 2):
        return None
    if quality is None:
        base_scores = [0] * L
    else:
        base_scores = [java_round(F(100) * PROB_CORRECT[qv]) - 100 for qv in quality]
    base = base_key_hit_score // 8
    rng = base_key_hit_score - base
    all_scores = [base + java_round(F(rng) * (F(1) - p)) for p in prob]
    key_scores = [all_scores[o] for o in offsets]
    all_err = F(1)
    for o in offsets:
        all_err = all_err * prob[o]
    if all_err > F(0.50):
        return None
    keys = [to_number(bases, o, o + k - 1) for o in offsets]
    n = len(offsets)
    keys_m = [rcomp_binary(keys[n - 1 - i], k) for i in range(n)]
    offsets_m = [L - (offsets[n - 1 - i] + k) for i in range(n)]
    return dict(offsets=offsets, keys=keys, keyScores=key_scores, baseScores=base_scores, offsetsM=offsets_m, keysM=keys_m)
