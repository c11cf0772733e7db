"""CPU: the index wire formats — bbmap_b200/csrc/wire.cpp through the C ABI against (a) the byte layout the Java Object
Serialization Stream Protocol prescribes, written out by hand here, (b) tests/javaser.py, an independent generic parser/writer of the protocol,
(c) the reference's own conventions: Block.write's delta-coded starts (align2/Block.java:104-118), the .gz rule (ReadWrite.java:372-381), the
file names of IndexMaker4.fname (:477-488), ChromosomeArray's serializable fields (dna/ChromosomeArray.java:415-419), summary.txt keys
(dna/FastaToChromArrays2.java:229-250).  No device needed."""
import gzip
import os
import struct

import numpy as np
import pytest

from bbmap_b200 import lib as _lib, wire

import javaser

INT_ARRAY_123 = bytes.fromhex("aced0005" "75" "72" "0002" "5b49" "4dba602676eab2a5" "02" "0000" "78" "70" "00000003" "00000001" "00000002" "00000003")


def test_int_array_bytes_are_what_java_writes(tmp_path):
    p = str(tmp_path / "a.block")
    wire.write_int_array(p, [1, 2, 3])
    assert open(p, "rb").read() == INT_ARRAY_123                      # ObjectOutputStream.writeObject(new int[]{1,2,3})
    assert wire.read_int_array(p).tolist() == [1, 2, 3]
    wire.write_int_array(p, [])
    assert open(p, "rb").read() == INT_ARRAY_123[:23] + b"\0\0\0\0" and wire.read_int_array(p).size == 0
    neg = np.array([-1, -2**31, 2**31 - 1, 0x01020304], np.int32)
    wire.write_int_array(p, neg)
    assert open(p, "rb").read()[27:] == struct.pack(">4i", *neg.tolist())      # big-endian two's complement
    assert np.array_equal(wire.read_int_array(p), neg)


def test_gz_suffix_selects_gzip(tmp_path):
    p = str(tmp_path / "x.block2.gz")
    wire.write_int_array(p, np.arange(1000))
    raw = open(p, "rb").read()
    assert raw[:2] == b"\x1f\x8b" and gzip.decompress(raw) == javaser.int_array_stream(list(range(1000)))
    assert np.array_equal(wire.read_int_array(p), np.arange(1000))
    # a gzip stream written by someone else (java.util.zip.GZIPOutputStream, `gzip -c`, pigz) reads the same
    q = str(tmp_path / "y.gz")
    open(q, "wb").write(gzip.compress(javaser.int_array_stream([7, 8, 9]), 9))
    assert wire.read_int_array(q).tolist() == [7, 8, 9]


def test_independent_parser_reads_what_the_library_writes(tmp_path):
    rng = np.random.default_rng(3)
    a = rng.integers(-2**31, 2**31 - 1, size=5000, dtype=np.int64).astype(np.int32)
    p = str(tmp_path / "r.bin")
    wire.write_int_array(p, a)
    o = javaser.parse_file(p)
    assert o["class"] == "[I" and o["suid"] == 0x4DBA602676EAB2A5 and o["values"] == a.tolist()


def test_block_files(tmp_path):
    k = 4; nk = 4 ** k
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 6, size=nk)
    starts = np.zeros(nk + 1, np.int32); np.cumsum(lens, out=starts[1:])
    sites = rng.integers(0, 2**30, size=int(starts[-1]), dtype=np.int64).astype(np.int32)
    fname = wire.block_fname(str(tmp_path) + "/ref/index/", 1, 3, 13, 2, 1)
    assert fname == str(tmp_path) + "/ref/index/1/chr1-3_index_k13_c2_b1.block"
    assert wire.block_fname("ref/index/", 4, 4, 13, 2, 7) == "ref/index/7/chr4_index_k13_c2_b7.block"
    wire.write_block(fname, sites, starts)
    # sites: raw object stream in <fname>; starts: gzip of the object stream of x[0]=starts[0], x[i]=starts[i]-starts[i-1] in <fname>2.gz
    assert javaser.parse_file(fname)["values"] == sites.tolist()
    d = javaser.parse_file(fname + "2.gz")["values"]
    assert d[0] == 0 and d[1:] == lens.tolist() and open(fname + "2.gz", "rb").read()[:2] == b"\x1f\x8b"
    s2, t2 = wire.read_block(fname)
    assert np.array_equal(s2, sites) and np.array_equal(t2, starts)
    # a Block written by the other implementation
    f2 = str(tmp_path / "other.block")
    open(f2, "wb").write(javaser.int_array_stream(sites.tolist()))
    open(f2 + "2.gz", "wb").write(gzip.compress(javaser.int_array_stream([0] + lens.tolist())))
    s3, t3 = wire.read_block(f2)
    assert np.array_equal(s3, sites) and np.array_equal(t3, starts)


def test_block_rejects_what_block_java_asserts(tmp_path):
    f = str(tmp_path / "bad.block")
    open(f, "wb").write(javaser.int_array_stream([1, 2, 3]))
    open(f + "2.gz", "wb").write(gzip.compress(javaser.int_array_stream([0, 1, 1, 1])))           # numStarts = 3: not a power of two (Block.java:35)
    with pytest.raises(_lib.BbmError, match="not a Block"):
        wire.read_block(f)
    open(f + "2.gz", "wb").write(gzip.compress(javaser.int_array_stream([0, 1, 1, 1, 1])))        # starts[numStarts] = 4 != sites.length
    with pytest.raises(_lib.BbmError, match="not a Block"):
        wire.read_block(f)


def test_bad_streams_fail_loudly(tmp_path):
    p = str(tmp_path / "t.bin")
    open(p, "wb").write(INT_ARRAY_123[:-2])
    with pytest.raises(_lib.BbmError, match="truncated"):
        wire.read_int_array(p)
    open(p, "wb").write(b"\xca\xfe" + INT_ARRAY_123[2:])
    with pytest.raises(_lib.BbmError, match="ACED0005"):
        wire.read_int_array(p)
    open(p, "wb").write(INT_ARRAY_123.replace(b"[I", b"[J"))
    with pytest.raises(_lib.BbmError, match=r"array class is \[J"):
        wire.read_int_array(p)
    with pytest.raises(_lib.BbmError, match="cannot open"):
        wire.read_int_array(str(tmp_path / "missing.block"))


def test_chromosome_array(tmp_path):
    arr = np.frombuffer(b"NNNNACGTACGTNNNN", np.uint8)
    p = str(tmp_path / "ref/genome/1/chr1.chrom.gz")
    wire.write_chrom(p, 1, arr)
    raw = gzip.decompress(open(p, "rb").read())
    assert raw == javaser.chromosome_array_stream(1, arr.tobytes(), 0, len(arr) - 1, 0)          # byte-for-byte the layout default serialization produces
    o = javaser.parse_file(p)
    assert o["class"] == "dna.ChromosomeArray" and o["suid"] == 3199182397853127842
    assert o["order"] == ["chromosome", "maxIndex", "minIndex", "strand", "array"]              # primitives by name, then object fields
    assert o["fields"]["array"]["class"] == "[B" and bytes(x & 255 for x in o["fields"]["array"]["values"]) == arr.tobytes()
    ca = wire.read_chrom(p)
    assert ca["chromosome"] == 1 and ca["minIndex"] == 0 and ca["maxIndex"] == 15 and ca["strand"] == 0 and ca["array"].tobytes() == arr.tobytes()
    q = str(tmp_path / "c7.chrom.gz")
    open(q, "wb").write(gzip.compress(javaser.chromosome_array_stream(7, b"ACGTN", 0, 4, 0)))
    ca = wire.read_chrom(q)
    assert ca["chromosome"] == 7 and ca["array"].tobytes() == b"ACGTN" and ca["maxIndex"] == 4


def test_genome_directory_round_trip(tmp_path):
    from bbmap_b200 import workloads as wl
    from bbmap_b200.index import pack_chromosomes
    scafs = [wl.random_genome(5000, seed=1), wl.random_genome(3000, seed=2)]
    cb, co, table = pack_chromosomes(scafs)
    root = str(tmp_path / "ref/genome")
    wire.write_genome(root, 1, cb, co, table, name="two_scaffolds")
    txt = open(os.path.join(root, "1", "summary.txt")).read().splitlines()
    assert txt[0] == "#Summary" and "chroms\t1" in txt and "bases\t%d" % len(cb) in txt and "defined\t8000" in txt and "interpad\t300" in txt and "name\ttwo_scaffolds" in txt
    cb2, co2, s = wire.read_genome(root, 1)
    assert np.array_equal(cb2, cb) and np.array_equal(co2, co) and s["scaffolds"] == 2 and s["defined"] == 8000 and s["undefined"] == len(cb) - 8000
