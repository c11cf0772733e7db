"""The C restatement of findTipDeletions / quickRescue (oracle/rescue_oracle.c) against hand-built cases whose answer follows from the
reference's text (current/align2/AbstractMapThread.java:1107-1141, 2178-2294, 2303-2405), and against an independent brute-force
formulation of the same rules.  No Java-generated vectors exist (no JVM): parity unpinned against Java itself."""
import numpy as np

from bbmap_b200 import rescue as rs
from rescue_cases import rescue_cases, tipdel_cases


def _b(s):
    return np.frombuffer(s.encode(), np.int8)


REF = "N" * 20 + "ACGTTGCAAGCTTAGGCTAACCGGTTAACGATCGATTACAGGATCCATGCAAGTCTGAACGTTAGCCTAGGATCAATCGGCTAAGCTTGCATGCCTGCAGG" + \
      "TCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTCACTGGCCGTCGTTTTACAACGTCGTGACTGGGAAAACCCTGGCGTTACCCAACTTAATCGCC" + "N" * 20


def test_tip_deletion_right_hand_case(oracle):
    ref = _b(REF)
    body = REF[30:70]                     # 40 bases placed at 30..69
    tip = REF[80:88]                      # the last 8 bases really come from 10 further right (a 10-base deletion)
    read = _b(body + tip)                 # 48 bases; ungapped site = 30..77
    # tip compared with REF[70:78]: count by hand what the reference counts
    orig = sum(1 for i in range(8) if (body + tip)[47 - i] != REF[77 - i])
    assert orig >= 3
    x = oracle.find_tip_deletions_right(read, ref, 20, 77, 100, 8)
    assert x == 10                        # start=87 reproduces the tip exactly: bestStart-originalStop = 10
    # a search range that ends before start=87 cannot return 10
    assert oracle.find_tip_deletions_right(read, ref, 20, 77, 9, 8) != 10
    # a read that matches its site has < 3 tip mismatches: 0
    assert oracle.find_tip_deletions_right(_b(REF[30:78]), ref, 20, 77, 100, 8) == 0
    # originalStop too close to minIndex: fail
    assert oracle.find_tip_deletions_right(read, ref, 75, 77, 100, 8) == 0


def test_tip_deletion_left_hand_case(oracle):
    ref = _b(REF)
    tip = REF[40:48]                      # first 8 bases come from 12 to the left of where the body puts them
    body = REF[60:100]
    read = _b(tip + body)                 # body says start = 52
    orig = sum(1 for i in range(8) if (tip + body)[i] != REF[52 + i])
    assert orig >= 3
    y = oracle.find_tip_deletions_left(read, ref, 20, 52, 100, 8)
    assert y == 12
    assert oracle.find_tip_deletions_left(_b(REF[52:100]), ref, 20, 52, 100, 8) == 0
    assert oracle.find_tip_deletions_left(read, ref, 52, 52, 100, 8) == 0      # minIndex >= originalStart


def test_quick_rescue_hand_cases(oracle):
    ref = _b(REF)
    L = 40
    read = REF[100:140]
    t = np.zeros(4, rs.RESCUE_TASK_DTYPE)
    #          read_off ref_off len ref_len       min max              loc search ideal mm right pad
    t[0] = (0, 0, L, len(ref), 20, len(ref) - 21, 60, 100, 95, 5, 1, 0)        # exact copy inside the range
    t[1] = (L, 0, L, len(ref), 20, len(ref) - 21, 60, 100, 95, 5, 1, 0)        # one substitution at read position 7
    t[2] = (0, 0, L, len(ref), 20, len(ref) - 21, 60, 30, 95, 5, 1, 0)         # range ends before the copy: null
    t[3] = (0, 0, L, len(ref), 20, len(ref) - 21, 150, 100, 95, 5, 0, 0)       # searching leftwards from 150 finds it too
    r1 = list(read); r1[7] = "A" if r1[7] != "A" else "C"
    reads = np.concatenate([_b(read), _b("".join(r1))])
    o = oracle.rescue_batch(reads, ref, t, rs.rescue_cfg())
    assert (o["start"][0], o["stop"][0], o["mismatches"][0], o["max_contig"][0]) == (100, 139, 0, 0)
    assert o["score"][0] == 70 + 100 * (L - 1) and o["perfect"][0] == 3 and o["in_bounds"][0] == 1
    assert (o["start"][1], o["mismatches"][1], o["max_contig"][1]) == (100, 1, 7)      # contig = run before the mismatch (:2339-2343)
    assert o["score"][1] == 70 + 100 * (L - 2) and o["perfect"][1] == 0
    assert o["start"][2] == -1
    assert (o["start"][3], o["mismatches"][3]) == (100, 0)


def _brute_tip(bases, ref, mn, orig, dist, tiplen, right):
    """Order-free statement: the first start in scan order with the minimum full mismatch count, if it beats the original."""
    L = len(bases)
    if right:
        if orig < mn + tiplen - 1 or orig >= len(ref):
            return 0
        cmpo = lambda s, j: bases[L - 1 - j] != ref[s - j]
    else:
        if orig + tiplen >= len(ref) or mn >= orig:
            return 0
        cmpo = lambda s, j: bases[j] != ref[s + j]
    om = last = contig = 0
    for i in range(tiplen):
        if contig >= 5:
            break
        if cmpo(orig, i):
            om += 1; last = i; contig = 0
        else:
            contig += 1
    if om < 3:
        return 0
    tl = last + 1
    if tl < 4:
        return 0
    if right:
        starts = range(orig + 1, min(len(ref) - 1, orig + min(dist, 30 * om)) + 1)
    else:
        starts = range(orig - 1, max(mn, orig - min(dist, 16 + 16 * om + 8 * tl)) - 1, -1)
    best = (om, None)
    for s in starts:
        m = sum(1 for j in range(tl) if cmpo(s, j))
        if m < best[0]:
            best = (m, s)
    if best[1] is None or best[0] > 2 or om - best[0] < 2:
        return 0
    return abs(best[1] - orig)


def test_tipdel_restatement_vs_bruteforce(oracle):
    g, reads, tasks = tipdel_cases(n=1500, seed=3)
    cfg = rs.tipdel_cfg()
    o = oracle.tipdel_batch(reads, g, tasks, cfg)
    changed = 0
    for i in range(len(tasks)):
        T = tasks[i]; L = int(T["read_len"]); b = reads[T["read_off"]:T["read_off"] + L]
        start, stop = int(T["start"]), int(T["stop"]); x = y = 0
        go = T["slow_score"] < T["max_imperfect"] and L > 16
        ms = min(100, 3000 - (8 + 8 + max(L, stop - start))) if go else 0
        if go and ms >= 1:
            if T["flags"] & 1:
                x = _brute_tip(b, g, int(T["min_index"]), stop, ms, 8, True)
                if x > 0:
                    stop += x; ms = min(ms, 3000 - (16 + max(L, stop - start)))
            if ms >= 1 and T["flags"] & 2:
                y = _brute_tip(b, g, int(T["min_index"]), start, ms, 8, False)
                if y > 0:
                    start -= y
        assert (o["start"][i], o["stop"][i], o["right"][i], o["left"][i]) == (start, stop, x, y), i
        changed += (x > 0) + (y > 0)
    assert changed > 200


def test_rescue_restatement_properties(oracle):
    g, reads, tasks = rescue_cases(n=1200, seed=4)
    o = oracle.rescue_batch(reads, g, tasks, rs.rescue_cfg())
    found = o["start"] >= 0
    assert found.sum() > 400 and (~found).sum() > 100
    for i in np.nonzero(found)[0]:
        T = tasks[i]; L = int(T["read_len"]); b = reads[T["read_off"]:T["read_off"] + L]; s = int(o["start"][i])
        w = g[s:s + L]
        mism = int(((b != w) | (b == ord("N"))).sum())
        assert mism == o["mismatches"][i] <= T["max_mismatches"] + 1
        lo, hi = (max(T["min_index"], T["loc"]), min(len(g) - L, T["loc"] + T["search_dist"])) if T["flags"] & 1 else \
                 (max(T["min_index"], T["loc"] - T["search_dist"]), min(len(g) - L, T["loc"]))
        assert lo <= s <= hi and o["stop"][i] == s + L - 1
        assert o["score"][i] == 70 + 100 * (L - 1 - mism)
        assert bool(o["perfect"][i] & 1) == (mism == 0)
    assert (tasks["read_len"][~found] < 10).any()


def _rescue_all_starts(b, g, T):
    """quickRescue stated over ALL candidate starts at once (numpy: full mismatch counts and the longest match run that a mismatch ends, for every
    start), followed by the reference's selection rule; the early exits of the reference's loops only skip candidates this rule rejects anyway."""
    L = len(b)
    if L < 10:
        return None
    right = bool(T["flags"] & 1)
    loc, dist, ideal = int(T["loc"]), int(T["search_dist"]), int(T["ideal_start"])
    lo, hi = (max(int(T["min_index"]), loc), min(len(g) - L, loc + dist)) if right else (max(int(T["min_index"]), loc - dist), min(len(g) - L, loc))
    if hi < lo:
        return None
    win = np.lib.stride_tricks.sliding_window_view(g[lo: hi + L], L)
    eq = (win == b[None, :]) & (b != ord("N"))[None, :]
    mism = L - eq.sum(axis=1)
    cur = np.zeros(len(win), np.int64); contig = np.zeros(len(win), np.int64)
    for j in range(L):
        contig = np.where(eq[:, j], contig, np.maximum(contig, cur))
        cur = np.where(eq[:, j], cur + 1, 0)
    score = (L - mism) + contig
    limit = int(T["max_mismatches"]) + 1
    first = lo
    best_start, best_m, best_c, best_score, best_ad = -1, 0, 0, 0, 1 << 40
    s = lo if right else hi
    while lo <= s <= hi:
        k = s - first
        m, sc, ad = int(mism[k]), int(score[k]), abs(s - ideal)
        if m <= limit and (sc > best_score or (sc == best_score and ad < best_ad)):
            best_start, best_m, best_c, best_score, best_ad = s, m, int(contig[k]), sc, ad
            limit = m
            if m == 0:
                if right:
                    hi = min(hi, ideal + ad)
                else:
                    lo = max(lo, ideal - ad)
        s += 1 if right else -1
    best = None if best_start < 0 else (best_start, best_m, best_c, best_score, best_ad)
    return best


def test_rescue_restatement_vs_all_starts_formulation(oracle):
    g, reads, tasks = rescue_cases(n=1500, seed=44)
    o = oracle.rescue_batch(reads, g, tasks, rs.rescue_cfg())
    found = shrunk = 0
    for i in range(len(tasks)):
        T = tasks[i]; L = int(T["read_len"]); b = reads[T["read_off"]:T["read_off"] + L]
        best = _rescue_all_starts(b, g, T)
        if best is None:
            assert o["start"][i] == -1, i
            continue
        found += 1
        s, m, c, sc, ad = best
        assert (o["start"][i], o["stop"][i], o["mismatches"][i], o["max_contig"][i]) == (s, s + L - 1, m, c), (i, best, o[i])
        assert o["score"][i] == 70 + 100 * (L - 1 - m)
        w = g[s:s + L]
        perfect = bool((b == w).all() and not (b == ord("N")).any())
        assert bool(o["perfect"][i] & 1) == perfect
    assert found > 500
