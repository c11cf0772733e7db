"""GPU parity of the whole chain (bench/pipeline.py): ingest -> seeds -> index search -> SiteScore lists -> removeOutOfBounds -> trimList ->
scoreNoIndels -> findTipDeletions -> scoreSlow (rounds, gapped sites included) -> final policy, everything resident on the device, against the
same chain through the CPU restatements: every field of every final site and the read flags, on a genome with several scaffolds."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bench"))


@pytest.mark.parametrize("genome_len,scaffolds,pairs,floor", [(400_000, 1, 4000, 0.99), (900_000, 5, 3000, 0.95)])   # the second genome carries 100-copy repeats
def test_chain_identical_to_cpu_chain(genome_len, scaffolds, pairs, floor):
    import pipeline
    res = pipeline.run(pairs=pairs, genome_len=genome_len, reps=1, device=0, cpu=True, scaffolds=scaffolds)
    cs = res["cpu_baseline"]
    assert cs["reads"] == 2 * pairs and cs["device_chain_identical_on_sample"] and cs["device_sam_fields_identical_on_sample"]
    assert res["mean_mapq"] > 20
    assert res["mapped"] > floor and res["top_site_is_origin"] > floor - 0.03 and res["status_nonzero"] == 0
    assert res["slow_alignments"] == cs["slow_alignments"] > pairs
