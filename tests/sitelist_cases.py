"""Seeded site lists for the list-policy tests (oracle and GPU)."""
import numpy as np

from bbmap_b200 import sitelist as sl
from bbmap_b200 import workloads as wl


def random_lists(nreads=4000, cap=40, seed=51, after_alignment=False):
    """Lists as they look after BBIndex.find (score = quickScore) or, with after_alignment, after scoreSlow (score = slowScore)."""
    rng = np.random.default_rng(seed)
    lists = np.zeros((nreads, cap), sl.SS_DTYPE); nss = np.zeros(nreads, np.int32)
    lens = rng.choice([50, 100, 150, 250], size=nreads)
    read_off = np.zeros(nreads + 1, np.int64); np.cumsum(lens, out=read_off[1:])
    for r in range(nreads):
        L = int(lens[r]); maxq = 70 + 100 * (L - 1)
        n = int(rng.choice([0, 1, 2, 3, 4, 5, 7, 9, 13, 17, 21, 26, 34, 39])) if rng.random() < 0.7 else int(rng.integers(0, cap + 1))
        n = min(n, cap)
        top = maxq if rng.random() < 0.3 else int(maxq * rng.uniform(0.3, 1.0))
        for i in range(n):
            s = lists[r, i]
            s["chrom"] = int(rng.integers(1, 4)); s["strand"] = int(rng.integers(0, 2))
            s["start"] = int(rng.integers(0, 5000)); s["stop"] = s["start"] + L - 1 + (int(rng.integers(0, 30)) if rng.random() < 0.2 else 0)
            s["hits"] = int(rng.integers(1, 18))
            sc = top if (i == 0 or rng.random() < 0.15) else int(top * rng.uniform(0.2, 1.0))
            if rng.random() < 0.2:
                sc = top - int(rng.integers(0, 500))
            sc = max(-200, min(sc, maxq))
            s["score"] = sc; s["quick_score"] = sc if not after_alignment else int(sc * rng.uniform(0.5, 1.0))
            if after_alignment:
                s["slow_score"] = sc
            s["perfect"] = 1 if sc == maxq and rng.random() < 0.8 else 0
            s["semiperfect"] = 1 if s["perfect"] or rng.random() < 0.05 else 0
            if rng.random() < 0.08:
                g = int(rng.choice([2, 4, 6])); s["ngaps"] = g; s["gaps"][:g] = np.sort(rng.integers(0, 9000, size=g))
        # duplicates: same place, possibly other gaps / scores
        for _ in range(int(rng.integers(0, 3))):
            if n >= 2 and n < cap:
                a = int(rng.integers(0, n)); lists[r, n] = lists[r, a]
                if rng.random() < 0.5:
                    lists[r, n]["score"] -= int(rng.integers(0, 300)); lists[r, n]["slow_score"] = lists[r, n]["score"] if after_alignment else 0
                if rng.random() < 0.3:
                    lists[r, n]["ngaps"] = 2; lists[r, n]["gaps"][:2] = (1, 2)
                if rng.random() < 0.3:
                    lists[r, n]["perfect"] = 0
                n += 1
        perm = rng.permutation(n); lists[r, :n] = lists[r, :n][perm]
        nss[r] = n
    return lists, nss, read_off


def noindel_lists(nreads=3000, cap=8, seed=61):
    """Reads cut from a genome with their true site plus decoys: shifted copies, sites longer than the read whose quick score claims a
    near-perfect hit (the stop-anchored rescoring of AbstractMapThread.java:806-813), sites hanging over the array ends, N blocks."""
    rng = np.random.default_rng(seed)
    g1 = wl.random_genome(30000, seed=seed).copy(); g2 = wl.random_genome(20000, seed=seed + 1).copy()
    g1[:300] = ord("N"); g1[-300:] = ord("N"); g2[:300] = ord("N"); g2[-300:] = ord("N"); g1[9000:9020] = ord("N")
    refs = np.concatenate([g1, g2]); chrom_off = np.array([0, len(g1), len(g1) + len(g2)], np.int64)
    lens = rng.choice([60, 100, 150], size=nreads)
    read_off = np.zeros(nreads + 1, np.int64); np.cumsum(lens, out=read_off[1:])
    P = np.zeros(int(read_off[-1]), np.int8); M = np.zeros_like(P)
    lists = np.zeros((nreads, cap), sl.SS_DTYPE); nss = np.zeros(nreads, np.int32)
    for r in range(nreads):
        L = int(lens[r]); ch = int(rng.integers(1, 3)); g = g1 if ch == 1 else g2
        p = int(rng.integers(320, len(g) - 320 - L))
        read = g[p:p + L].copy()
        read = np.where(read == ord("N"), ord("A"), read).astype(np.int8)
        for _ in range(int(rng.choice([0, 0, 0, 1, 2, 6]))):
            k = int(rng.integers(0, L)); read[k] = ord("ACGT"[("ACGT".index(chr(read[k])) + int(rng.integers(1, 4))) % 4])
        if rng.random() < 0.05:
            read[int(rng.integers(0, L))] = ord("N")
        strand = int(rng.integers(0, 2))
        plus = read if strand == 0 else wl.revcomp(read.view(np.uint8)).view(np.int8)
        P[read_off[r]:read_off[r + 1]] = plus; M[read_off[r]:read_off[r + 1]] = wl.revcomp(plus.view(np.uint8)).view(np.int8)
        maxq = 70 + 100 * (L - 1)
        n = int(rng.integers(0, cap + 1))
        for i in range(n):
            s = lists[r, i]; kind = int(rng.integers(0, 7))
            s["chrom"] = ch; s["strand"] = strand; s["start"] = p; s["stop"] = p + L - 1
            s["score"] = int(maxq * rng.uniform(0.4, 1.0)); s["hits"] = 5
            if kind == 1:
                d = int(rng.integers(-3, 4)); s["start"] += d; s["stop"] += d
            elif kind == 2:                    # longer site, read really sits at its end
                s["start"] = p - int(rng.integers(1, 40)); s["score"] = maxq - int(rng.integers(0, 400))
            elif kind == 3:                    # longer site, read sits at its start
                s["stop"] = p + L - 1 + int(rng.integers(1, 40)); s["score"] = maxq - int(rng.integers(0, 400))
            elif kind == 4:
                s["strand"] = 1 - strand
            elif kind == 5:
                s["start"] = int(rng.choice([-5, -1, len(g) - L + 2, len(g) - 3, 8990])); s["stop"] = s["start"] + L - 1
            elif kind == 6 and (read == g[p:p + L]).all() and strand == 0:
                s["perfect"] = 1; s["semiperfect"] = 1; s["score"] = maxq
            s["quick_score"] = s["score"]
            if rng.random() < 0.1:
                s["ngaps"] = 2; s["gaps"][:2] = (s["start"], s["stop"])
        nss[r] = n
    return refs, chrom_off, P, M, read_off, lists, nss


def slow_cases(nreads=1500, cap=4, seed=81):
    """Lists as they stand when processRead calls scoreSlow: every site scored by scoreNoIndels (slow_score = score), sorted.  Reads carry
    substitutions and one indel of 1-60 bp, some of them within 12 bases of a tip (the aligner then asks for more padding); decoy sites,
    sites longer than the read, sites at the ends of the array, a few gapped sites, and reads for which scoreSlow is not called."""
    rng = np.random.default_rng(seed)
    g = wl.random_genome(40000, seed=seed).copy()
    g[:200] = ord("N"); g[-200:] = ord("N")
    refs = g; chrom_off = np.array([0, len(g)], np.int64)
    lens = rng.choice([60, 100, 150, 250], size=nreads)
    read_off = np.zeros(nreads + 1, np.int64); np.cumsum(lens, out=read_off[1:])
    P = np.zeros(int(read_off[-1]), np.int8); M = np.zeros_like(P)
    lists = np.zeros((nreads, cap), sl.SS_DTYPE); nss = np.zeros(nreads, np.int32); run = np.ones(nreads, np.int32)
    for r in range(nreads):
        L = int(lens[r])
        p = int(rng.integers(260, len(g) - 260 - L - 70)) if rng.random() < 0.93 else int(rng.choice([198, 203, len(g) - 200 - L - 3, len(g) - 200 - L + 2]))
        kind = int(rng.integers(0, 6)); d = int(rng.integers(1, 61))
        q = int(rng.integers(3, 13)) if rng.random() < 0.35 else int(rng.integers(13, L - 13))
        if rng.random() < 0.5:
            q = L - q
        if kind in (0, 1):          # deletion of d bases after read position q
            src = np.concatenate([g[p:p + q], g[p + q + d:p + L + d]])
        elif kind == 2:             # insertion of min(d, 20) random bases at q
            d2 = min(d, 20, L - q - 1)
            src = np.concatenate([g[p:p + q], wl.ACGT[rng.integers(0, 4, size=d2, dtype=np.uint8)].view(np.int8), g[p + q:p + L - d2]])
        else:
            src = g[p:p + L]
        read = np.where(src == ord("N"), ord("A"), src).astype(np.int8)[:L]
        for _ in range(int(rng.choice([0, 1, 2, 4]))):
            k = int(rng.integers(0, L)); read[k] = ord("ACGT"[("ACGT".index(chr(read[k])) + int(rng.integers(1, 4))) % 4])
        strand = int(rng.integers(0, 2))
        plus = read if strand == 0 else wl.revcomp(read.view(np.uint8)).view(np.int8)
        P[read_off[r]:read_off[r + 1]] = plus; M[read_off[r]:read_off[r + 1]] = wl.revcomp(plus.view(np.uint8)).view(np.int8)
        n = int(rng.integers(1, cap + 1))
        for i in range(n):
            s = lists[r, i]
            s["chrom"] = 1; s["strand"] = strand; s["start"] = p; s["stop"] = p + L - 1; s["hits"] = 4
            if i > 0:
                w = int(rng.integers(0, 5))
                if w == 0:
                    s["start"] = int(rng.integers(220, len(g) - 300 - L)); s["stop"] = s["start"] + L - 1          # decoy
                elif w == 1:
                    s["stop"] = s["start"] + L - 1 + d                                                               # site spanning the deletion
                elif w == 2:
                    s["strand"] = 1 - strand
                elif w == 3:
                    s["start"] += int(rng.integers(-6, 7)); s["stop"] = s["start"] + L - 1
                elif s["start"] + L + 300 < len(g) - 210:      # gapped site (kept inside the array: quickMap's removeOutOfBounds guarantees that)
                    s["ngaps"] = 4; s["stop"] = s["start"] + L + 300; s["gaps"][:4] = (s["start"], s["start"] + 40, s["start"] + 340, s["stop"])
            s["quick_score"] = int((70 + 100 * (L - 1)) * rng.uniform(0.4, 1.0)); s["score"] = s["quick_score"]
        nss[r] = n
        if rng.random() < 0.1:
            run[r] = 0
    return refs, chrom_off, P, M, read_off, lists, nss, run


def random_match_strings(read_off, seed=7):
    """Long-format match strings ('m', 'S', 'N', 'I', 'D', 'C', rarely 'X'/'Y') consuming exactly each read, with events concentrated in the tips;
    every 17th read has none (r.match == null), every 41st is cut short inside the tip, every 53rd is short format."""
    rng = np.random.default_rng(seed)
    n = len(read_off) - 1; parts = []; off = np.zeros(n + 1, np.int64)
    for r in range(n):
        L = int(read_off[r + 1] - read_off[r]); s = []
        pos = 0
        while pos < L:
            tip = pos < 9 or pos >= L - 9
            u = rng.random()
            if u < (0.25 if tip else 0.02):
                k = rng.random()
                if k < 0.35:
                    s.append("S"); pos += 1
                elif k < 0.5:
                    s.append("N"); pos += 1
                elif k < 0.6:
                    s.append("C"); pos += 1
                elif k < 0.75:
                    s.append("I"); pos += 1
                elif k < 0.97:
                    s.append("D" * int(rng.integers(1, 5)))
                else:
                    s.append("X" if pos < 9 else "Y"); pos += 1
            else:
                s.append("m"); pos += 1
        t = "".join(s)
        if r % 17 == 16:
            t = ""
        elif r % 41 == 40:
            t = t[:5]
        elif r % 53 == 52:
            t = "m3S" + t[5:]
        parts.append(t); off[r + 1] = off[r] + len(t)
    return np.frombuffer("".join(parts).encode(), np.int8).copy(), off


def homopolymer_reads(read_off, seed=9):
    """Read bases (as sequenced) with homopolymer runs and Ns at the tips."""
    rng = np.random.default_rng(seed)
    b = np.frombuffer(b"ACGT", np.int8)[rng.integers(0, 4, size=int(read_off[-1]))].copy()
    for r in range(len(read_off) - 1):
        a, e = int(read_off[r]), int(read_off[r + 1])
        if rng.random() < 0.4:
            b[a:a + int(rng.integers(2, 10))] = b[a]
        if rng.random() < 0.4:
            b[e - int(rng.integers(2, 10)):e] = b[e - 1]
        if rng.random() < 0.05:
            b[a] = ord("N"); b[a + 1] = ord("N")
    return b
