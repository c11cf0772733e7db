"""CPU: the control flow of BBMapThread.scoreSlow — the sequential C restatement the CUDA rounds are checked against (oracle/scoreslow_oracle.c) must
equal a second restatement written from the Java text alone (tests/pyscoreslow.py), whose alignments are MSA.fillAndScoreLimited restated in tests/pygapped.py with
every fill done by the reference's own C: every site field after scoreSlow (gap arrays included), gapped sites too."""
import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from sitelist_cases import slow_cases

import pyscoreslow
from test_sitelist_independent import _same, _to_sites


@pytest.mark.parametrize("seed,kw", [(505, {}), (506, dict(paired=1, min_ratio_pre_rescue=0.336))])
def test_score_slow_control_flow(oracle, seed, kw):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=260, seed=seed)
    pcfg = sl.policy_cfg()
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, pcfg, P, M, refs, co)
    scfg = sl.slow_cfg(**kw)
    L2, status, na = oracle.score_slow(lists, nss, ro, P, M, refs, co, run, scfg)
    packed = oracle.new_packed(601, 3000)
    P8 = np.ascontiguousarray(P).view(np.int8); M8 = np.ascontiguousarray(M).view(np.int8); R8 = np.ascontiguousarray(refs).view(np.int8)
    checked = fills = retried = gapped = 0
    for r in range(len(nss)):
        n = int(nss[r])
        if not run[r] or n == 0 or status[r]:
            continue
        sites = _to_sites(lists[r], n)
        chroms = {s.chrom for s in sites}
        if len(chroms) != 1:
            continue
        ch = chroms.pop()
        ref8 = R8[int(co[ch - 1]): int(co[ch])]
        a, b = int(ro[r]), int(ro[r + 1])
        before = [(s.start, s.stop) for s in sites]
        fills += pyscoreslow.score_slow(oracle, packed, sites, P8[a:b], M8[a:b], ref8, scfg[0])
        _same(sites, L2[r], n, r)
        retried += sum(1 for s, (x, y) in zip(sites, before) if (s.stop - s.start) != (y - x))
        checked += 1; gapped += int((lists[r, :n]["ngaps"] > 0).any())
    assert checked > 120 and fills > 150 and retried > 30 and gapped > 10, (checked, fills, retried, gapped)
