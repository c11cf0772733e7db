"""CPU: the chunked (32 read positions at a time) formulations that walk_warp_kernel (bbmap_b200/csrc/search_walk_warp.cuh) uses for extendScore and
MSA.calcAffineScore, restated in Python lane by lane — ballots as lists, prefix population counts as sums — against the sequential restatements of
tests/pyfind.py on thousands of extensions from a repeat-rich genome (keys pointing at several near-identical copies, indels, N) and on random
locArrays.  A formulation slip (e.g. a stop after the 32nd position of a chunk, which a first version missed) fails here without a GPU."""
import numpy as np

from bbmap_b200 import workloads as wl
from bbmap_b200.index import pack_chromosomes
from bbmap_b200.keyring import default_cfg

import pyfind


def warp_affine(loc, bs):
    L = len(loc); score = 0; carry_last_loc = -3; carry_last_value = -1; carry_run = 0
    for base in range(0, L, 32):
        ch = [loc[base + j] if base + j < L else 0 for j in range(32)]
        inn = [base + j < L for j in range(32)]
        pos = [inn[j] and ch[j] > 0 for j in range(32)]; sub = [inn[j] and ch[j] == -1 for j in range(32)]
        for j in range(32):
            last_value = ch[j - 1] if j > 0 else carry_last_value
            below = [k for k in range(j) if pos[k]]
            last_loc = ch[below[-1]] if below else carry_last_loc
            if pos[j]:
                l = ch[j]; b = bs[base + j]
                if l == last_value:
                    score += 100 + b
                elif l == last_loc or last_loc < 0:
                    score += 70 + b
                elif l < last_loc:
                    dif = last_loc - l + 1; s = 0
                    if dif > 256:
                        s += ((dif - 128) // 128) * -2; dif = dif % 128 + 128
                    if dif > 80:
                        s += ((dif - 80 + 3) // 4) * -1; dif = 80
                    if dif > 20:
                        s += (dif - 20) * -1; dif = 20
                    if dif > 5:
                        s += (dif - 5) * -9; dif = 5
                    if dif > 1:
                        s += (dif - 1) * -33
                    score += 70 + b - 472 + s
                else:
                    score += 70 + b - 395 - 39 * (min(l - last_loc, 5) - 1)
            elif sub[j]:
                nz = [k for k in range(j) if not sub[k]]
                run = j - nz[-1] if nz else j + 1 + carry_run
                score += -25 if run > 5 else (-51 if run > 1 else -127)
        pp = [k for k in range(32) if pos[k]]
        if pp:
            carry_last_loc = ch[pp[-1]]
        carry_last_value = ch[31]
        if sub[31]:
            nz = [k for k in range(32) if not sub[k]]
            carry_run = 31 - nz[-1] if nz else 32 + carry_run
        else:
            carry_run = 0
    return score


def warp_extend(py, bases, offsets, values, chrom, center, nh):
    K = py.KEYLEN; L = len(bases); loc = [-1] * L
    center_val = values[center]; center_loc = py.number_to_site(center_val)
    ref = py.chroms[chrom]

    def chunk(clocs, refbase, misses, post_rule):
        lanes = []
        for cloc in clocs:
            rloc = refbase + cloc
            valid = 0 <= cloc < L and 0 <= rloc < len(ref)
            lanes.append((cloc, valid, loc[cloc] if valid else 0, valid and bases[cloc] == ref[rloc]))
        mm = [v and not m for (_c, v, _o, m) in lanes]
        pre = [(not v) or o == refbase or (misses + sum(mm[:j]) > 0 and o >= 0) for j, (_c, v, o, _m) in enumerate(lanes)]
        post = [v and (not m) and post_rule(o) for (_c, v, o, m) in lanes]
        fpre = pre.index(True) if True in pre else 32
        fpost = post.index(True) if True in post else 32
        nproc = min(fpre, fpost + 1)
        for j, (c, _v, o, m) in enumerate(lanes):
            if j < nproc and m and (o < 0 or refbase == center_loc):
                loc[c] = refbase
        return misses + sum(mm[:nproc]), (True in pre) or (True in post)

    keynum = 0
    for i in range(nh):
        if not (center_val - pyfind.MAX_INDEL <= values[i] <= center_val + pyfind.MAX_INDEL2):
            continue
        refbase = py.number_to_site(values[i]); keynum += 1; misses = 0
        top = offsets[i] + K - 1
        while top >= 0:
            misses, stop = chunk([top - j for j in range(32)], refbase, misses, lambda o, kn=keynum: o >= 0 or kn > 1)
            if stop:
                break
            top -= 32
    for i in range(nh):
        if not (center_val - pyfind.MAX_INDEL <= values[i] <= center_val + pyfind.MAX_INDEL2):
            continue
        refbase = py.number_to_site(values[i]); misses = 0
        bot = offsets[i] + K
        while bot < L:
            misses, stop = chunk([bot + j for j in range(32)], refbase, misses, lambda o: o >= 0)
            if stop:
                break
            bot += 32
    for i in range(L):
        if bases[i] == ord("N"):
            loc[i] = -2
    return loc


def test_affine_scan_on_random_loc_arrays():
    rng = np.random.default_rng(1)
    for _ in range(3000):
        L = int(rng.integers(20, 301))
        loc = []
        cur = int(rng.integers(1000, 100000))
        while len(loc) < L:
            kind = rng.random(); run = int(rng.integers(1, 70))
            if kind < 0.55:
                loc += [cur] * run
            elif kind < 0.75:
                loc += [-1] * run
            elif kind < 0.8:
                loc += [-2] * int(rng.integers(1, 4))
            else:
                cur += int(rng.integers(-400, 400)) if kind < 0.95 else int(rng.integers(-30000, 30000))
                cur = max(cur, 1)
        loc = loc[:L]
        bs = rng.integers(-40, 1, size=L).tolist()
        assert warp_affine(loc, bs) == pyfind.calc_affine_score(list(loc), bs)


def test_chunked_extension_on_repeat_families(oracle):
    rng = np.random.Generator(np.random.PCG64(77))
    g = wl.random_genome(200_000, seed=76)
    starts = []
    for _ in range(6):
        unit = wl.ACGT[rng.integers(0, 4, size=400, dtype=np.uint8)]
        for _c in range(5):
            u = unit.copy(); m = rng.random(400) < 0.012
            u[m] = wl.ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
            q = int(rng.integers(0, len(g) - 400)); g[q:q + 400] = u; starts.append(q)
    cb, co, table = pack_chromosomes([g])
    n_ext = 0
    for L in (100, 150, 250):
        R = wl.make_mapping_reads(cb, co, table, 200, L=L, seed=80 + L, sub_rate=0.02, indel_rate=0.01)
        bases, qual, off = R["bases"].copy(), R["qual"], R["off"]
        for i in range(0, len(off) - 1, 2):
            q = min(max(starts[int(rng.integers(0, len(starts)))] + int(rng.integers(-L // 2, 400 - L // 2)), 0), len(g) - L)
            r = g[q:q + L].copy()
            if rng.random() < 0.3:
                r[int(rng.integers(0, L))] = ord("N")
            bases[off[i]:off[i + 1]] = r if (i // 2) % 2 == 0 else wl.revcomp(r)
        idx = oracle.index_build(cb, co, 13, -1)
        es = oracle.seed_batch(bases, qual, off, default_cfg(), 32)
        cfg, blocks, counts, hist = idx
        py = pyfind.BBIndexPy(cfg, blocks, counts, hist, cb, co, False)
        sequential = py.extend_score
        seen = []

        def both(b, bs, offsets, values, chrom, center, loc, nh):
            sc = sequential(b, bs, offsets, values, chrom, center, loc, nh)
            w = warp_extend(py, b, offsets, values, chrom, center, nh)
            assert w == list(loc), [(p, loc[p], w[p]) for p in range(len(loc)) if loc[p] != w[p]][:8]
            assert warp_affine(w, bs) == sc
            seen.append(1)
            return sc
        py.extend_score = both
        for i in range(len(off) - 1):
            nk = int(es["nkeys"][i])
            if nk >= 1:
                a, b = int(off[i]), int(off[i + 1])
                py.find(bases[a:b].tobytes(), es["baseScores"][a:b].view(np.int8), es["offsets"][i, :nk], es["keyScores"][i, :nk])
        n_ext += len(seen)
    assert n_ext > 1500, n_ext
