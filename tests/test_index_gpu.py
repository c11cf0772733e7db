"""GPU parity: device index build + analysis vs the C restatement of IndexMaker4 / BBIndex.analyzeIndex — starts, sites,
COUNTS, lengthHistogram and every derived limit, bit-exact; plus the structural invariants the reference relies on."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _genome(seed, sizes, repeats=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    scafs = []
    for n in sizes:
        s = wl.ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)].copy()
        if repeats and n > 3000:
            s[1000:1400] = ord("A")                       # homopolymer: banned keys + clumps
            s[1500:1900] = np.tile(np.frombuffer(b"AC", np.uint8), 200)
            s[2000:2600] = np.tile(np.frombuffer(b"ACG", np.uint8), 200)   # period 3: clumpy but not banned
            s[2700:2720] = ord("N")
            unit = s[100:400].copy()
            for r in range(5):                             # planted repeats -> lists longer than 1
                p = int(rng.integers(3000, n - 400)); s[p:p + 300] = unit
        scafs.append(s)
    return scafs


@pytest.mark.parametrize("k,sizes,chrombits", [(10, (60000, 5000, 80000), -1), (11, (30000,) * 5, 1), (13, (200000, 70000), -1), (9, (3000,), 0)])
def test_index_build_parity(oracle, k, sizes, chrombits):
    from bbmap_b200.index import BBIndexCUDA, pack_chromosomes
    scafs = _genome(100 + k, sizes)
    # force several chromosomes for the multi-block cases
    bytes_, off, table = pack_chromosomes(scafs, max_length=120000 if len(sizes) > 1 else (1 << 29) - 200000)
    ecfg, eblocks, ecounts, ehist = oracle.index_build(bytes_, off, k, chrombits)
    idx = BBIndexCUDA(bytes_, off, keylen=k, chrombits=chrombits)
    try:
        assert idx.cfg.tobytes() == ecfg.tobytes(), (idx.cfg, ecfg)
        assert idx.nblocks == len(eblocks)
        for b, (es, et) in enumerate(eblocks):
            gs, gt, gc, gh = idx.download(b)
            assert np.array_equal(gs, es) and np.array_equal(gt, et)
            # lists sorted by (chrom,pos)
            if len(gt) > 1:
                inner = np.ones(len(gt), bool); inner[gs[:-1][gs[:-1] < len(gt)]] = False
                assert (np.diff(gt.astype(np.int64))[inner[1:]] > 0).all()
        assert np.array_equal(gc, ecounts) and np.array_equal(gh, ehist)
        # banned keys (period <= 2) never indexed; COUNTS symmetric under reverse complement
        assert gc[0] == 0 and gc[(1 << (2 * k)) - 1] == 0
        assert (gc > 0).any()
    finally:
        idx.close()


@pytest.mark.parametrize("k,sizes,chrombits", [(10, (60000, 5000, 80000), -1), (11, (30000,) * 5, 1)])
def test_index_save_load_in_reference_format(oracle, tmp_path, k, sizes, chrombits):
    """the resident index written the way IndexMaker4 writes it (Block.write) and loaded back instead of built; the files are checked
    with the independent parser (tests/javaser.py) against the oracle's arrays, and the loaded index (+ its recomputed analysis) equals the built one."""
    import os
    import javaser
    from bbmap_b200 import wire
    from bbmap_b200.index import BBIndexCUDA, pack_chromosomes
    scafs = _genome(100 + k, sizes)
    bytes_, off, table = pack_chromosomes(scafs, max_length=120000)
    ecfg, eblocks, ecounts, ehist = oracle.index_build(bytes_, off, k, chrombits)
    root = str(tmp_path) + "/ref/index/"
    idx = BBIndexCUDA(bytes_, off, keylen=k, chrombits=chrombits)
    try:
        idx.save(root, build=1)
        cbits = int(idx.cfg[0]["chrombits"]); cpb = 1 << cbits; nch = len(off) - 1
        names = sorted(os.listdir(root + "1"))
        assert len(names) == 2 * idx.nblocks
        b = 0; chrom = 1
        while chrom <= nch:
            lo = max(1, chrom & ~(cpb - 1)); hi = min(nch, (chrom & ~(cpb - 1)) + cpb - 1)
            fname = wire.block_fname(root, lo, hi, k, cbits, 1)
            es, et = eblocks[b]
            assert javaser.parse_file(fname)["values"] == et.tolist()
            delta = javaser.parse_file(fname + "2.gz")["values"]
            assert delta[0] == int(es[0]) and delta[1:] == np.diff(es).tolist()
            b += 1; chrom = hi + 1
        assert b == idx.nblocks
    finally:
        idx.close()
    idx2 = BBIndexCUDA(bytes_, off, keylen=k, chrombits=chrombits, load_from=root, build=1)
    try:
        assert idx2.cfg.tobytes() == ecfg.tobytes() and idx2.nblocks == len(eblocks)
        for b, (es, et) in enumerate(eblocks):
            gs, gt, gc, gh = idx2.download(b)
            assert np.array_equal(gs, es) and np.array_equal(gt, et)
        assert np.array_equal(gc, ecounts) and np.array_equal(gh, ehist)
    finally:
        idx2.close()
    with pytest.raises(Exception, match="cannot open"):
        BBIndexCUDA(bytes_, off, keylen=k, chrombits=chrombits, load_from=root, build=2)
