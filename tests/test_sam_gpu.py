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
