"""Seeded cases for the tip-deletion search and the mate-rescue scan (shared by the oracle tests and the GPU parity tests)."""
import numpy as np

from bbmap_b200 import rescue as rs
from bbmap_b200 import workloads as wl

MIN_INDEX = 700          # leading 'N' padding of the chromosome array (ChromosomeArray.minIndex)


def make_genome(n=60000, seed=5):
    g = wl.random_genome(n, seed=seed).copy()
    g[:MIN_INDEX] = ord("N"); g[-900:] = ord("N")
    g[20000:20040] = ord("N")
    return g, MIN_INDEX, n - 901           # array, minIndex, maxIndex


def _mutate(rng, read, nsub, nN=0):
    read = read.copy()
    for _ in range(nsub):
        k = int(rng.integers(0, len(read)))
        read[k] = ord("ACGT"[(("ACGT".index(chr(read[k])) if chr(read[k]) in "ACGT" else 0) + int(rng.integers(1, 4))) % 4])
    for _ in range(nN):
        read[int(rng.integers(0, len(read)))] = ord("N")
    return read


def tipdel_cases(n=6000, seed=31):
    rng = np.random.default_rng(seed)
    g, mn, mx = make_genome()
    reads, tasks = [], np.zeros(n, rs.TIPDEL_TASK_DTYPE)
    off = 0
    for i in range(n):
        L = int(rng.choice([12, 16, 17, 40, 100, 150, 250]))
        kind = int(rng.integers(0, 8))
        p = int(rng.integers(mn + 300, mx - 700))
        t = int(rng.integers(3, 13)); d = int(rng.integers(1, 160))
        start, stop = p, p + L - 1
        if kind in (0, 1) and L > t:        # deletion of d bases t from the right tip
            read = np.concatenate([g[p:p + L - t], g[p + L - t + d:p + L + d]])
        elif kind in (2, 3) and L > t:      # deletion of d bases t from the left tip; the body is placed correctly
            read = np.concatenate([g[p - d:p - d + t], g[p + t:p + L]])
        elif kind == 4:                     # both tips
            t2 = int(rng.integers(3, 9)); d2 = int(rng.integers(1, 90))
            if L > t + t2:
                read = np.concatenate([g[p - d2:p - d2 + t2], g[p + t2:p + L - t], g[p + L - t + d:p + L + d]])
            else:
                read = g[p:p + L].copy()
        elif kind == 5:                     # unrelated read
            q = int(rng.integers(mn, mx - L)); read = g[q:q + L].copy()
        elif kind == 6:                     # sites at the ends of the array / next to the N block
            start = int(rng.choice([mn - 3, mn, mn + 2, mn + 7, mx - L - 5, mx - L + 1, len(g) - L - 1, len(g) - L, 20040 - L + 5, 19990]))
            stop = start + L - 1
            read = _mutate(rng, g[max(0, start):max(0, start) + L], 0)
            read = np.where(read == ord("N"), ord("A"), read).astype(np.int8)
            read[-6:] = [ord(c) for c in "TGCATG"][:min(6, L)][-6:] if L >= 6 else read[-6:]
        else:
            read = g[p:p + L].copy()
        read = _mutate(rng, np.asarray(read, np.int8), int(rng.integers(0, 3)), nN=int(rng.integers(0, 2)) if rng.random() < 0.1 else 0)
        if len(read) != L:
            read = np.resize(read, L)
        if rng.random() < 0.05:
            stop = start + L - 1 + int(rng.integers(1, 3200))          # an already-extended site: maxSearch shrinks or vanishes
        reads.append(read)
        slow = int(rng.integers(0, 100 * L))
        tasks[i] = (off, 0, L, len(g), mn, start, stop, slow, slow + 1 if rng.random() < 0.85 else slow, int(rng.integers(0, 4)))
        off += L
    return g, np.concatenate(reads).astype(np.int8), tasks


def rescue_cases(n=3000, seed=37):
    rng = np.random.default_rng(seed)
    g, mn, mx = make_genome(seed=6)
    g = g.copy()
    reads, tasks = [], np.zeros(n, rs.RESCUE_TASK_DTYPE)
    off = 0
    # planted exact repeats so that several perfect / near-perfect placements fall inside one search range
    for k in range(40):
        a = int(rng.integers(mn + 2000, mx - 4000)); b = a + int(rng.integers(160, 900))
        g[b:b + 260] = g[a:a + 260]
    for i in range(n):
        L = int(rng.choice([8, 10, 36, 100, 150, 250]))
        p = int(rng.integers(mn + 1500, mx - 2500))
        kind = int(rng.integers(0, 6))
        read = g[p:p + L].copy()
        if kind == 1:
            read = _mutate(rng, read, int(rng.integers(1, 6)))
        elif kind == 2:
            read = _mutate(rng, read, int(rng.integers(5, 40)), nN=int(rng.integers(0, 3)))
        elif kind == 3:
            q = int(rng.integers(mn, mx - L)); read = g[q:q + L].copy()       # unrelated: usually null
        elif kind == 4:
            read = _mutate(rng, read, 0, nN=int(rng.integers(1, 4)))
        right = int(rng.integers(0, 2))
        dist = int(rng.choice([0, 50, 400, 1200, 1500]))
        if right:
            loc = p - int(rng.integers(0, dist + 40))
        else:
            loc = p + int(rng.integers(0, dist + 40))
        if kind == 5:                       # ranges clipped by the ends of the array
            loc = int(rng.choice([mn - 500, mn + 10, mx - 100, len(g) - 50]))
            read = g[max(mn, min(loc, mx - L)):][:L].copy()
            read = np.where(read == ord("N"), ord("C"), read).astype(np.int8)
        ideal = p + int(rng.integers(-300, 300))
        mm = int(rng.choice([0, 1, 5, 32, max(0, int(0.6 * L - 1))]))
        reads.append(np.asarray(read, np.int8))
        tasks[i] = (off, 0, L, len(g), mn, mx, loc, dist, ideal, mm, right, 0)
        off += L
    return g, np.concatenate(reads).astype(np.int8), tasks
