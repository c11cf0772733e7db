"""CPU: scoreNoIndels and the gapped-reference path (a15: makeGref, coordinate translation, the dispatch rule and the -120 of the
Java wrapper, score / traceback on a gapped reference) — the C restatement the CUDA kernels are checked against (oracle/host_oracle.c,
oracle/msa_oracle.c) must equal a second restatement written from the Java text alone (tests/pygapped.py), with every fill done by the reference's
own C and every walk by tests/pywalk.py."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

import pygapped


def test_score_no_indels_independent(oracle):
    genome = wl.random_genome(30000, seed=226).copy()
    genome[100:130] = ord("N")
    reads, tasks = wl.make_msa_tasks(genome, 1500, seed=227, flags=0, n_rate=0.01)
    nt = np.zeros(len(tasks) + 60, wl.NOINDEL_TASK_DTYPE)
    n0 = len(tasks)
    nt["read_off"][:n0] = tasks["read_off"]; nt["read_len"][:n0] = tasks["read_len"]
    nt["ref_len"] = len(genome); nt["ref_start"][:n0] = tasks["ref_start"] + 4
    for k in range(60):                      # sites hanging over both ends of the chromosome array, N blocks
        j = n0 + k
        nt["read_off"][j] = tasks["read_off"][k]; nt["read_len"][j] = tasks["read_len"][k]
        nt["ref_start"][j] = [-7, -1, len(genome) - 30, len(genome) - 1, 90, 120][k % 6]
    g = genome.view(np.int8).tolist()
    r8 = reads.view(np.int8)
    for flags in (0, 1):
        nt["flags"] = flags
        moff = np.zeros(len(nt) + 1, np.int64); np.cumsum(nt["read_len"], out=moff[1:])
        exp, em = oracle.noindel_batch(reads, genome, nt, match_off=moff if flags else None)
        subs = oob = 0
        for i, t in enumerate(nt):
            rd = r8[int(t["read_off"]): int(t["read_off"]) + int(t["read_len"])].tolist()
            if flags:
                s, m = pygapped.score_no_indels(rd, g, int(t["ref_start"]), True)
                assert s == exp[i], (i, s, exp[i])
                if m is None:
                    oob += 1
                else:
                    assert bytes(m) == em[moff[i]: moff[i + 1]].tobytes(), i
                    subs += b"S" in bytes(m)
            else:
                assert pygapped.score_no_indels(rd, g, int(t["ref_start"])) == exp[i], i
        if flags:
            assert subs > 300 and oob >= 30


def _spliced_case(rng, genome, L):
    nex = int(rng.integers(2, 4))
    cuts = np.sort(rng.choice(np.arange(25, L - 25), size=nex - 1, replace=False))
    lens = np.diff(np.concatenate([[0], cuts, [L]]))
    pos = int(rng.integers(9000, len(genome) - 12000))
    exons, gaps, p = [], [], pos
    for ln in lens:
        exons.append(genome[p:p + ln].copy()); gaps += [p, p + ln - 1]
        p += ln + int(rng.integers(300, 2500))
    read = np.concatenate(exons)
    for q in rng.integers(0, L, size=int(rng.integers(0, 3))):
        read[q] = wl.ACGT[rng.integers(0, 4)]
    if rng.random() < 0.3:                                  # a short deletion inside an exon
        q = int(rng.integers(10, L - 10)); read = np.concatenate([read[:q], read[q + 2:], read[:2]])
    return read, np.array(gaps, np.int32)


def test_gapped_reference_independent(oracle):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    rng = np.random.Generator(np.random.PCG64(321))
    genome = wl.random_genome(60000, seed=308)
    g8 = genome.view(np.int8)
    MAXR, MAXC = 601, 3000
    packed = oracle.new_packed(MAXR, MAXC)
    gapped_ok = unlimited = failed = padded = 0
    for i in range(150):
        L = int(rng.choice([100, 150, 250]))
        if i % 4 == 3:
            p = int(rng.integers(9000, 40000)); read = genome[p:p + L].copy(); g = None; lo, hi = p, p + L - 1
            if i % 8 == 7:
                read = np.concatenate([read[:40], read[45:], read[:5]])       # 5-base deletion
        else:
            read, g = _spliced_case(rng, genome, L); lo, hi = int(g[0]), int(g[-1])
        pad = int(rng.integers(0, 12))
        ms = [int(0.3 * wl.max_quality(L)), int(0.95 * wl.max_quality(L)), 0][i % 3 if i % 5 else 2]
        a, b = lo - pad, hi + pad
        sc, match, max4 = oracle.fill_and_score_limited_gapped(read, genome, a, b, ms, g)
        sv, mstr, m4 = pygapped.fill_and_score_limited(oracle, packed, MAXR, MAXC, read.view(np.int8), g8, a, b, ms, g)
        if sc is None:
            assert sv is None, i
            failed += 1
            continue
        assert sv == sc, (i, sv, sc)
        assert m4 == max4.tolist(), (i, m4, max4)
        assert mstr == match.tobytes(), (i, mstr, match.tobytes())
        padded += len(sv) == 8
        if g is not None:
            gapped_ok += 1
            assert mstr.count(b"D") >= 128
        unlimited += ms == 0
    assert gapped_ok >= 80 and unlimited >= 20 and failed >= 5, (gapped_ok, unlimited, failed, padded)


def test_gref_layout_by_hand():
    """One intron of 1000 bases: 64 + 1000 % 128 = 168 intron bases kept on the left, (1000 - 128) / 128 = 6 gap symbols, 64 kept on the right."""
    ref = list(range(0, 3000))                               # value == position (mod 256 irrelevant here: a plain list)
    g = pygapped.Gref(ref, [100, 199, 1200, 1299], 96, 1303)
    assert g.origin == 96 and g.bytes[0] == 96
    body = g.bytes[: g.limit]
    assert body[:104 + 168] == list(range(96, 200 + 168))
    assert body[272:278] == [pygapped.GAPC] * 6
    assert body[278:] == list(range(1136, 1304))
    assert g.limit == 104 + 168 + 6 + 64 + 104 and g.limit2 == g.limit + 127
    # an intron base count of 1000 = 168 + 6 * 128 + 64: coordinates on both sides of the gap translate back exactly
    for p in (96, 150, 367, 1136, 1200, 1303):
        assert g.from_gapped(g.to_gapped(p)) == p
    assert g.to_gapped(1136) == 278 and g.from_gapped(272) == 368 and g.from_gapped(273) == 368 + 128


def test_score_of_a_match_string_independent(oracle):
    """MSA.score(match) (what realign_new and the tip-penalty code re-score match strings with): groupby formulation vs the C restatement."""
    import ctypes as C
    rng = np.random.default_rng(77)
    syms = np.frombuffer(b"mmmmmmmmmmmmSSNIDDXYCR", np.uint8)
    cases = [b"m" * 100, b"m" * 40 + b"S" + b"m" * 59, b"mS" + b"m" * 98, b"N" + b"S" * 7 + b"m" * 20, b"m" * 30 + b"D" * 300 + b"m" * 70, b"m" * 30 + b"D" * 257 + b"m" * 70,
             b"m" * 30 + b"D" * 256 + b"m" * 70, b"X" * 3 + b"m" * 50 + b"I" * 25 + b"m" * 50 + b"Y" * 2, b"C" * 5 + b"m" * 95, b"m", b"S", b"D" * 16000 + b"m"]
    for _ in range(3000):
        n = int(rng.integers(1, 300))
        runs = []
        while sum(len(x) for x in runs) < n:
            s = syms[rng.integers(0, len(syms))]
            runs.append(bytes([s]) * int(rng.choice([1, 1, 2, 3, 6, 21, 81, 90, 300]) if s in b"DIS" else rng.integers(1, 40)))
        cases.append(b"".join(runs))
    oracle.lib.orc_score_match.restype = C.c_int
    for m in cases:
        b = np.frombuffer(m, np.int8)
        exp = oracle.lib.orc_score_match(b.ctypes.data_as(C.c_void_p), C.c_int(len(b)))
        assert pygapped.score_match(m) == exp, (m[:80], pygapped.score_match(m), exp)
    assert pygapped.score_match(b"m" * 100) == 9970 and pygapped.score_match(b"m" * 30 + b"DDD" + b"m" * 70) == 2970 + 6970 - 472 - 66
