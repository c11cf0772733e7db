"""CPU: the SiteScore helpers of realign_new / genMatchString — leftPaddingNeeded, rightPaddingNeeded, clipTipIndels (+ unclip, MSA.score, setPerfect),
fixXY, setPerfect of the C restatement (oracle/mapper_oracle.c, through its test entry point) must equal a second restatement written from the Java text alone
(tests/pyclip.py) on the edited match string, start / stop, the three scores, the perfect bits and the return value."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from bbmap_b200 import workloads as wl

import pyclip
import pysitelist as ps


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _case(rng, genome, kind):
    """A site with a match string that agrees with its start/stop, the read it describes and — for the X/Y cases — tips that hang over the alignment."""
    L = int(rng.choice([40, 100, 150]))
    start = int(rng.integers(300, len(genome) - 800)) if rng.random() < 0.9 else int(rng.choice([0, 3, len(genome) - L - 2, len(genome) - L - 40]))
    body = []
    syms = "mmmmmmmmmmmmmmmmmmmmSN"
    while sum(1 for c in body if c != "D") < L:
        r = rng.random()
        if kind == "clip" and r < 0.04:
            body += ["I"] * int(rng.choice([1, 2, 4, 12])) if rng.random() < 0.5 else ["D"] * int(rng.choice([1, 3, 9, 45]))
        elif r < 0.01:
            body += ["I"] * int(rng.integers(1, 3)) if rng.random() < 0.5 else ["D"] * int(rng.integers(1, 4))
        else:
            body.append(syms[int(rng.integers(0, len(syms)))])
    # trim to exactly L read symbols
    out, used = [], 0
    for c in body:
        if c == "D" or used < L:
            out.append(c); used += c != "D"
    while out and out[-1] == "D":
        out.pop()
    if kind == "clip" and rng.random() < 0.6:            # an indel burst inside the tip
        k = int(rng.integers(0, 4))
        burst = ["I"] * int(rng.choice([2, 4, 11, 14])) if rng.random() < 0.5 else ["D"] * int(rng.choice([5, 13, 41, 50]))
        if rng.random() < 0.5:
            out = out[:k] + burst + out[k:]
        else:
            out = out[:len(out) - k] + burst + out[len(out) - k:]
    if kind == "xy":
        nx = int(rng.choice([0, 1, 3, 7, 12])); ny = int(rng.choice([0, 1, 2, 6, 11]))
        for i in range(min(nx, len(out))):
            out[i] = "X" if rng.random() < 0.8 else "Y"
        for i in range(min(ny, len(out))):
            out[len(out) - 1 - i] = "Y" if rng.random() < 0.8 else "X"
    match = "".join(out).encode()
    L = sum(1 for c in out if c != "D")
    reflen = sum(1 for c in out if c != "I")
    # X/Y symbols do not move the mapped interval in a site that still has them (fixXY extends it): half of the cases
    stop = start + reflen - 1
    if kind == "xy" and rng.random() < 0.5:
        lead = len(match) - len(match.lstrip(b"XY")); trail = len(match) - len(match.rstrip(b"XY"))
        start += lead; stop -= trail
        if stop < start:
            stop = start
    # the read: reference bases along the alignment, random elsewhere, a few N
    bases = np.empty(L, np.int8); rloc, cloc = start, 0
    for c in out:
        if c == "D":
            rloc += 1; continue
        g = genome[rloc] if (c != "I" and 0 <= rloc < len(genome)) else wl.ACGT[rng.integers(0, 4)]
        if c == "S":
            g = wl.ACGT[(int(np.searchsorted(wl.ACGT, min(g, ord("T")))) + 1) % 4]
        if c == "N" or rng.random() < 0.01:
            g = ord("N")
        if kind == "xy" and c in "XY" and rng.random() < 0.5:
            g = wl.ACGT[rng.integers(0, 4)]
        bases[cloc] = g; cloc += 1
        if c != "I":
            rloc += 1
    maxq = 70 + 100 * (L - 1)
    slow = int(maxq * rng.uniform(0.3, 1.0))
    return start, stop, match, bases, slow


@pytest.mark.parametrize("kind,op,tiplen,max_indel", [("clip", 2, 4, 10), ("clip", 2, 8, 3), ("clip", 2, 4, 1), ("clip", 0, 4, 10), ("clip", 1, 4, 10), ("clip", 0, 8, 2),
                                                      ("clip", 1, 8, 2), ("xy", 3, 0, 0), ("xy", 0, 4, 10), ("xy", 1, 4, 10), ("plain", 5, 0, 0), ("clip", 4, 0, 0)])
def test_site_helpers(oracle, kind, op, tiplen, max_indel):
    rng = np.random.default_rng(400 + op * 10 + tiplen + max_indel)
    genome = wl.random_genome(20000, seed=41).copy()
    genome[5000:5030] = ord("N")
    g8 = genome.view(np.int8); glist = g8.tolist()
    co = np.array([0, len(genome)], np.int64)
    lib = oracle.lib
    lib.orc_test_site_op.restype = C.c_int
    changed = nonzero = 0
    for it in range(1500):
        start, stop, match, bases, slow = _case(rng, genome, kind if kind != "plain" else "clip")
        if kind == "plain":                                   # setPerfect: ungapped sites on (nearly) matching reads, N in read or reference, ends of the array
            L = len(bases); stop = start + L - 1 + (1 if rng.random() < 0.05 else 0)
            bases = g8[start:start + L].copy() if start + L <= len(genome) else bases
            for _ in range(int(rng.choice([0, 0, 1, 3]))):
                bases[int(rng.integers(0, L))] = ord("N") if rng.random() < 0.5 else wl.ACGT[rng.integers(0, 4)]
            if rng.random() < 0.1:
                start = 5000 - int(rng.integers(0, L)); stop = start + L - 1
                bases = np.where(g8[start:start + L] == ord("N"), ord("A"), g8[start:start + L]).astype(np.int8)
        if op == 4:                                           # unclip: turn some tip symbols into C first
            mm = bytearray(match); k = int(rng.integers(0, 6))
            for i in range(k):
                if chr(mm[i]) in "mSN":
                    mm[i] = ord("C")
            for i in range(int(rng.integers(0, 6))):
                if chr(mm[len(mm) - 1 - i]) in "mSN":
                    mm[len(mm) - 1 - i] = ord("C")
            match = bytes(mm)
        rec = np.zeros(1, sl.SS_DTYPE)
        rec["chrom"] = 1; rec["start"] = start; rec["stop"] = stop; rec["score"] = slow + 7; rec["slow_score"] = slow; rec["quick_score"] = slow // 2
        rec["paired_score"] = slow + 120 if rng.random() < 0.3 else 0; rec["perfect"] = int(rng.integers(0, 2)); rec["semiperfect"] = 1
        site = ps.Site(1, 0, start, stop, 0, slow + 7, slow // 2, slow, int(rec["paired_score"][0]), bool(rec["perfect"][0]), True, False, None)
        cs = pyclip.ClipSite(site, match)
        mbuf = np.zeros(len(match) + 64, np.int8); mbuf[:len(match)] = np.frombuffer(match, np.int8)
        mlen = np.array([len(match)], np.int32)
        exp = lib.orc_test_site_op(C.c_int(op), _p(rec), _p(mbuf), _p(mlen), C.c_int(len(mbuf)), _p(bases), C.c_int(len(bases)), _p(g8), _p(co), C.c_int(tiplen), C.c_int(max_indel))
        bl = bases.tolist()
        if op == 0:
            got = cs.left_padding_needed(tiplen, max_indel)
        elif op == 1:
            got = cs.right_padding_needed(tiplen, max_indel)
        elif op == 2:
            got = int(cs.clip_tip_indels(bl, glist, tiplen, max_indel))
        elif op == 3:
            got = int(cs.fix_xy(bl, glist))
        elif op == 4:
            got = int(cs.unclip(bl, pyclip._getter(glist)))
        else:
            got = int(pyclip.set_perfect(site, bl, glist))
        assert got == exp, (it, op, got, exp, match, start, stop)
        r = rec[0]
        assert bytes(cs.match) == mbuf[: int(mlen[0])].tobytes(), (it, bytes(cs.match), mbuf[: int(mlen[0])].tobytes(), match)
        assert (site.start, site.stop, site.score, site.slowScore, site.pairedScore, int(site.perfect), int(site.semiperfect)) == \
               (int(r["start"]), int(r["stop"]), int(r["score"]), int(r["slow_score"]), int(r["paired_score"]), int(r["perfect"]), int(r["semiperfect"])), (it, match)
        changed += bytes(cs.match) != match or site.start != start or site.stop != stop
        nonzero += got != 0
    if op in (2, 3, 4):
        assert changed > 150, changed
    assert nonzero > 50
