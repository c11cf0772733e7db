"""GPU parity: SAM record fields on the device vs the C restatement — FLAG, POS, MAPQ, RNAME/RNEXT, PNEXT, TLEN and the CIGAR text,
bit-exact, both SAM versions, soft clipping on/off, an intron limit, paired and unpaired records."""
import numpy as np
import pytest

from bbmap_b200 import sam
from sam_cases import make_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("version,soft_clip,intron", [(1.4, 1, 2 ** 31 - 1), (1.3, 1, 2 ** 31 - 1), (1.4, 0, 2), (1.3, 0, 1)])
def test_sam_parity(oracle, version, soft_clip, intron):
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    tasks, mbuf, scaf = make_cases(n=20000, seed=11)
    cfg = sam.default_cfg(version); cfg["soft_clip"] = soft_clip; cfg["intron_limit"] = intron
    exp, ecb, coff = oracle.sam_batch(tasks, mbuf, scaf, cfg)
    m = MultiStateAligner11tsCUDA(device=0)
    try:
        got, gcb, goff = sam.sam_batch(m.h, tasks, mbuf, scaf, cfg)
    finally:
        m.close()
    assert np.array_equal(goff, coff)
    for f in exp.dtype.names:
        assert np.array_equal(got[f], exp[f]), f
    for i in range(len(tasks)):
        n = exp["cigar_len"][i]
        if n > 0:
            assert gcb[coff[i]:coff[i] + n].tobytes() == ecb[coff[i]:coff[i] + n].tobytes(), i
    assert (exp["cigar_len"] > 0).sum() > 15000 and (exp["flag"] & 0x2).any() and (exp["tlen"] < 0).any()


def test_tasks_from_lists_parity():
    """Read.setFromTopSite on the device (bbm_sam_tasks_from_lists_dev) vs its numpy statement: every field of every record, mapped and
    unmapped reads, ambiguous reads, with and without match strings."""
    import ctypes as C

    import torch
    from bbmap_b200 import lib as _lib
    from bbmap_b200 import sitelist as sl
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    from sitelist_cases import random_lists, random_match_strings
    lists, nss, ro = random_lists(nreads=5000, cap=6, seed=31, after_alignment=True)
    rng = np.random.default_rng(4)
    fl = np.zeros(len(nss), sl.READ_OUT_DTYPE)
    fl["flags"] = np.where(nss > 0, 1, 0) | np.where(rng.random(len(nss)) < 0.2, 4, 0)
    fl["flags"][::13] &= ~1                                                     # mapping cleared although the list is not empty
    _, mo = random_match_strings(ro, 3)
    L = _lib.load(); dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)
    q = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    m = MultiStateAligner11tsCUDA(device=0)
    try:
        for moff in (None, mo):
            d_t = torch.full((len(nss) * sam.SAM_TASK_DTYPE.itemsize,), 0x5a, dtype=torch.uint8, device=dev)
            d_l, d_n, d_o, d_f = up(lists), up(nss.astype(np.int32)), up(ro.astype(np.int64)), up(fl)
            d_m = None if moff is None else up(moff.astype(np.int64))
            _lib.check(L.bbm_sam_tasks_from_lists_dev(m.h, q(d_l), q(d_n), len(nss), lists.shape[1], q(d_o), q(d_f), q(d_m), q(d_t), None), "tasks_from_lists")
            torch.cuda.synchronize()
            got = np.frombuffer(d_t.cpu().numpy().tobytes(), sam.SAM_TASK_DTYPE)
            exp = sam.tasks_from_lists(lists, nss, ro, fl, moff)
            for f in exp.dtype.names:
                assert np.array_equal(got[f], exp[f]), f
        assert (exp["flags"] & sam.RF_MAPPED).any() and (exp["chrom"] == -1).any() and (exp["match_len"] > 0).any()
    finally:
        m.close()
