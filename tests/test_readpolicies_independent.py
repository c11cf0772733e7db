"""CPU: scoreNoIndels(Read, ...) and findTipDeletions(Read, ...) — the C restatement the CUDA kernels are checked against (oracle/sitelist_oracle.c)
must equal a second restatement written from the Java text (tests/pyreadpolicies.py) on every site field afterwards and on the per-read return values."""
import numpy as np
import pytest

from bbmap_b200 import rescue as rs
from bbmap_b200 import sitelist as sl
from sitelist_cases import noindel_lists, slow_cases

import pyreadpolicies as prp
from test_sitelist_independent import _same, _to_sites


def _refs(refs, co):
    r8 = np.ascontiguousarray(refs).view(np.int8)
    return {c: r8[int(co[c - 1]): int(co[c])].tolist() for c in range(1, len(co))}


def test_score_no_indels_read(oracle):
    refs, co, P, M, ro, lists, nss = noindel_lists(nreads=700, seed=907)
    L2, n2, out = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, sl.policy_cfg(), P, M, refs, co)
    R = _refs(refs, co)
    P8 = np.ascontiguousarray(P).view(np.int8); M8 = np.ascontiguousarray(M).view(np.int8)
    near = forced = moved = 0
    for r in range(len(nss)):
        n = int(nss[r])
        sites = _to_sites(lists[r], n)
        a, b = int(ro[r]), int(ro[r + 1])
        ret = prp.score_no_indels_read(sites, P8[a:b].tolist(), M8[a:b].tolist(), R)
        # the reference sorts the list right after (Collections.sort, BBMapThread.java:441): the C policy includes that sort
        import functools
        import pysitelist as ps
        sites.sort(key=functools.cmp_to_key(ps.compare_to))
        _same(sites, L2[r], int(n2[r]), r)
        assert ret == int(out["near_perfect"][r]), (r, ret, out[r])
        near += ret > 0; forced += ret < 0
        moved += sum(1 for s, t in zip(sorted(sites, key=lambda s: s.tag), lists[r, :n]) if s.start != int(t["start"]))
    assert near > 100 and forced > 10 and moved > 10, (near, forced, moved)


@pytest.mark.parametrize("seed,with_quality", [(981, True), (986, False)])
def test_find_tip_deletions_read(oracle, seed, with_quality):
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=900, seed=seed)
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, sl.policy_cfg(), P, M, refs, co)
    rng = np.random.default_rng(seed)
    quality = None
    if with_quality:
        quality = np.full(len(P), 30, np.int8)
        for r in rng.choice(len(nss), size=200, replace=False):
            a, b = int(ro[r]), int(ro[r + 1])
            if rng.random() < 0.5:
                quality[a:a + 8] = rng.integers(0, 16, size=8)
            if rng.random() < 0.5:
                quality[b - 8:b] = rng.integers(0, 16, size=8)
    R = _refs(refs, co)
    P8 = np.ascontiguousarray(P).view(np.int8); M8 = np.ascontiguousarray(M).view(np.int8)
    total = 0
    for cfg, mi in ((rs.tipdel_cfg(), None), (rs.tipdel_cfg(search_range=30), np.array([250] * (len(co) - 1), np.int32))):
        exp, eo = oracle.sitelist_tipdel(lists, nss, ro, P, M, quality, refs, co, cfg, mi)
        mins = {c: (0 if mi is None else int(mi[c - 1])) for c in range(1, len(co))}
        for r in range(len(nss)):
            n = int(nss[r])
            if (lists[r, :n]["ngaps"] > 0).any():
                continue
            sites = _to_sites(lists[r], n)
            a, b = int(ro[r]), int(ro[r + 1])
            ch = prp.find_tip_deletions_read(sites, P8[a:b].tolist(), M8[a:b].tolist(), None if quality is None else quality[a:b], R, mins,
                                             int(cfg["search_range"][0]), int(cfg["slow_rescue_padding"][0]))
            _same(sites, exp[r], n, r)
            assert ch == int(eo["best_sites"][r]), (r, ch, eo[r])
            total += ch
    assert total > 15, total
