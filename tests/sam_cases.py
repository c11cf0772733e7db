"""Shared generator for the SAM-record tests: reads mapped (or not) onto a packed multi-scaffold reference, with match strings made of
the symbols BBMap emits (m S N I D X Y C), paired and unpaired, reaching over scaffold ends and into the inter-scaffold padding."""
import numpy as np

from bbmap_b200 import sam


def make_cases(n=4000, seed=5, nchrom_scafs=((3000, 1200, 5000), (800,), (2500, 2500))):
    rng = np.random.Generator(np.random.PCG64(seed))
    table, pos = [], None
    for ch, scafs in enumerate(nchrom_scafs, start=1):
        p = 8000
        for ln in scafs:
            table.append((ch, p, ln)); p += ln + 300
    scaf = sam.scaffold_table(table, len(nchrom_scafs))
    tasks = np.zeros(n, sam.SAM_TASK_DTYPE)
    chunks, off = [], 0
    syms_mid = np.frombuffer(b"mmmmmmmmmmmmmmmmmmmmSSNIDDmmmm", np.uint8)
    for i in range(n):
        L = int(rng.choice([36, 100, 150, 250]))
        ch, st, ln = table[int(rng.integers(0, len(table)))]
        kind = rng.random()
        if kind < 0.35:
            m = np.full(L, ord("m"), np.uint8)
        else:
            m = syms_mid[rng.integers(0, len(syms_mid), size=L)].copy()
            if kind > 0.85:
                m[: int(rng.integers(1, 6))] = ord("C")
            if kind > 0.9:
                m[-int(rng.integers(1, 6)):] = ord("C")
            if 0.6 < kind < 0.65:
                m[0] = ord("X")
            if 0.65 < kind < 0.7:
                m[-1] = ord("Y")
            # make the string consume exactly L read bases (D consumes none): pad or trim in the middle
            while (m != ord("D")).sum() < L:
                m = np.insert(m, len(m) // 2, ord("m")).astype(np.uint8)
            while (m != ord("D")).sum() > L:
                k = len(m) // 2
                while m[k] == ord("D"):
                    k += 1
                m = np.delete(m, k)
        reflen = int((m != ord("I")).sum() - (m == ord("X")).sum() - (m == ord("Y")).sum())
        start = st + int(rng.integers(-12, max(ln - 10, 1)))
        if rng.random() < 0.05:
            start = st + ln - int(rng.integers(1, 30))          # over the end, may reach the next scaffold
        stop = start + max(reflen, 1) - 1
        flags = 0
        if rng.random() < 0.92:
            flags |= sam.RF_MAPPED
        if rng.random() < 0.5:
            flags |= sam.RF_MINUS
        if (m == ord("m")).all() and rng.random() < 0.9:
            flags |= sam.RF_PERFECT
        if rng.random() < 0.1:
            flags |= sam.RF_AMBIGUOUS
        if rng.random() < 0.03:
            flags |= sam.RF_SECONDARY
        if rng.random() < 0.02:
            flags |= sam.RF_DISCARDED
        has_match = rng.random() < 0.95
        maxq = 70 + 100 * (L - 1)
        tasks[i] = (off, len(m) if has_match else 0, ch, start, stop, L, int(rng.integers(int(0.4 * maxq), maxq + 1)), -1, flags, 0)
        chunks.append(m); off += len(m)
    # pair up a third of the records (neighbours), same chromosome half of the time
    for i in range(0, n - 1, 3):
        tasks["mate"][i] = i + 1; tasks["mate"][i + 1] = i
        tasks["flags"][i + 1] |= sam.RF_PAIRNUM1
        if rng.random() < 0.6:
            tasks["flags"][i] |= sam.RF_PAIRED; tasks["flags"][i + 1] |= sam.RF_PAIRED
        if rng.random() < 0.6:
            d = int(rng.integers(-400, 400))
            ln = tasks["stop"][i + 1] - tasks["start"][i + 1]
            tasks["chrom"][i + 1] = tasks["chrom"][i]; tasks["start"][i + 1] = tasks["start"][i] + d; tasks["stop"][i + 1] = tasks["start"][i + 1] + ln
    return tasks, np.concatenate(chunks), scaf


def cigar_stats(c):
    """(read bases consumed, reference bases consumed) of a CIGAR string."""
    import re
    q = r = 0
    for n, op in re.findall(r"(\d+)([MIDNSHP=X])", c):
        n = int(n)
        if op in "MIS=X":
            q += n
        if op in "MDN=X":
            r += n
    return q, r
