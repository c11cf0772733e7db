"""GPU: the single-alignment twins of the reference's C entry points and the Java_align2_* JNI symbols, driven through a
fake JNIEnv, leave the caller's `packed` matrix, result, iteration counter, vertLimit and horizLimit bit-identical to the
reference's own C (or the port) run on the same sequence of calls with a persistent matrix."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu
MAXR, MAXC = 601, 3000


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_jni_twins_bit_exact(oracle):
    from bbmap_b200 import lib as L
    lib = L.load()
    kind = "reference" if oracle.has_reference else "port"
    orc = oracle.lib
    fnL = C.cast(lib.Java_align2_MultiStateAligner11tsJNI_fillLimitedXJNI, C.c_void_p)
    fnU = C.cast(lib.Java_align2_MultiStateAligner11tsJNI_fillUnlimitedJNI, C.c_void_p)
    genome = wl.random_genome(30000, seed=61)
    reads, tasks = wl.make_msa_tasks(genome, 60, seed=62, flags=0)
    g8 = genome.view(np.int8)
    pa = oracle.new_packed(MAXR, MAXC); pb = pa.copy()
    vla = np.zeros(MAXR + 1, np.int32); hla = np.zeros(MAXC + 1, np.int32)
    vlb = vla.copy(); hlb = hla.copy()
    ita = np.zeros(1, np.int64); itb = 0
    n_fail = 0
    for i, t in enumerate(tasks):
        r = np.ascontiguousarray(reads[t["read_off"]: t["read_off"] + t["read_len"]].view(np.int8))
        bw, ratio = [(0, 0.0), (12, 0.0), (0, 0.2), (40, 0.0)][i % 4]
        ms = int(t["min_score"]) - 120
        res = np.zeros(5, np.int32)
        if i % 7 == 3:
            pins = orc.fake_call_fillUnlimitedJNI(fnU, _p(r), len(r), _p(g8), len(g8), int(t["ref_start"]), int(t["ref_end"]), _p(res), _p(ita),
                                                  _p(pa), len(pa), _p(oracle.sub), _p(oracle.ins), 604, MAXR, MAXC)
            exp, it = oracle.fill_unlimited(r, g8, int(t["ref_start"]), int(t["ref_end"]), pb, MAXR, MAXC, kind=kind)
            assert res[:4].tolist() == exp.tolist()
        else:
            pins = orc.fake_call_fillLimitedXJNI(fnL, _p(r), len(r), _p(g8), len(g8), int(t["ref_start"]), int(t["ref_end"]), ms, _p(res), _p(ita),
                                                 _p(pa), len(pa), _p(oracle.sub), _p(oracle.ins), 604, MAXR, MAXC, bw, C.c_float(ratio),
                                                 _p(vla), _p(hla), _p(oracle.b2n), _p(oracle.insC))
            exp, it = oracle.fill_limited(r, g8, int(t["ref_start"]), int(t["ref_end"]), ms, pb, MAXR, MAXC, bandwidth=bw, ratio=ratio,
                                          kind=kind, vl=vlb, hl=hlb)
            assert res.tolist() == exp.tolist(), (i, res, exp)
            n_fail += int(exp[4])
            assert np.array_equal(vla, vlb) and np.array_equal(hla, hlb)
        assert pins == 0
        itb += it
        assert int(ita[0]) == itb
        assert np.array_equal(pa, pb), "packed differs after call %d" % i
    assert 0 < n_fail < len(tasks)


def test_banded_jni_symbols(oracle):
    """The four Java_align2_BandedAlignerJNI_* entry points, through the fake JNIEnv, against the oracle."""
    from bbmap_b200 import lib as L
    lib = L.load()
    orc = oracle.lib
    orc.fake_call_bandedJNI.restype = C.c_int
    orc.fake_call_bandedRCJNI.restype = C.c_int
    names = ["alignForwardJNI", "alignForwardRCJNI", "alignReverseJNI", "alignReverseRCJNI"]
    fns = [C.cast(getattr(lib, "Java_align2_BandedAlignerJNI_" + n), C.c_void_p) for n in names]
    b2n = np.zeros(128, np.int8); comp = np.zeros(128, np.int8)
    orc.orc_banded_tables(_p(b2n), _p(comp))
    q, r, tasks = wl.make_banded_tasks(40, seed=71, min_len=30, max_len=400)
    for t in tasks:
        qq = np.ascontiguousarray(q[t["query_off"]: t["query_off"] + t["query_len"]].view(np.int8))
        rr = np.ascontiguousarray(r[t["ref_off"]: t["ref_off"] + t["ref_len"]].view(np.int8))
        rv = np.zeros(5, np.int32)
        d = int(t["dir"])
        if d in (0, 2):
            e = orc.fake_call_bandedJNI(fns[d], _p(qq), len(qq), _p(rr), len(rr), int(t["qstart"]), int(t["rstart"]), int(t["max_edits"]),
                                        C.c_ubyte(int(t["exact"])), int(t["max_width"]), _p(b2n), _p(rv))
        else:
            e = orc.fake_call_bandedRCJNI(fns[d], _p(qq), len(qq), _p(rr), len(rr), int(t["qstart"]), int(t["rstart"]), int(t["max_edits"]),
                                          C.c_ubyte(int(t["exact"])), int(t["max_width"]), _p(b2n), _p(comp), _p(rv))
        exp = oracle.banded(d, qq, rr, int(t["qstart"]), int(t["rstart"]), int(t["max_edits"]), bool(t["exact"]), int(t["max_width"]))
        assert (int(e), rv.tolist()) == exp


def test_jni_concurrent_threads(oracle):
    """Four host threads call Java_align2_MultiStateAligner11tsJNI_fillLimitedXJNI at the same time, each with its own arrays, the way BBMap's mapping
    threads do (one MSA per thread): every thread's sequence of calls must leave its `packed`, result and counters as the reference's C does."""
    import threading
    from bbmap_b200 import lib as L
    lib = L.load()
    kind = "reference" if oracle.has_reference else "port"
    orc = oracle.lib
    fnL = C.cast(lib.Java_align2_MultiStateAligner11tsJNI_fillLimitedXJNI, C.c_void_p)
    genome = wl.random_genome(30000, seed=71)
    g8 = genome.view(np.int8)
    reads, tasks = wl.make_msa_tasks(genome, 64, seed=72, flags=0)
    errors = []

    def worker(tid):
        try:
            pa = oracle.new_packed(MAXR, MAXC)
            vla = np.zeros(MAXR + 1, np.int32); hla = np.zeros(MAXC + 1, np.int32); ita = np.zeros(1, np.int64)
            out = []
            for t in tasks[tid::4]:
                r = np.ascontiguousarray(reads[t["read_off"]: t["read_off"] + t["read_len"]].view(np.int8))
                res = np.zeros(5, np.int32)
                orc.fake_call_fillLimitedXJNI(fnL, _p(r), len(r), _p(g8), len(g8), int(t["ref_start"]), int(t["ref_end"]), int(t["min_score"]) - 120, _p(res), _p(ita),
                                              _p(pa), len(pa), _p(oracle.sub), _p(oracle.ins), 604, MAXR, MAXC, 0, C.c_float(0.0),
                                              _p(vla), _p(hla), _p(oracle.b2n), _p(oracle.insC))
                out.append(res.copy())
            results[tid] = (out, pa, int(ita[0]))
        except Exception as e:          # noqa: BLE001
            errors.append(e)

    results = {}
    ths = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t_ in ths:
        t_.start()
    for t_ in ths:
        t_.join()
    assert not errors, errors
    for tid in range(4):
        pb = oracle.new_packed(MAXR, MAXC); itb = 0
        out, pa, ita = results[tid]
        for j, t in enumerate(tasks[tid::4]):
            r = np.ascontiguousarray(reads[t["read_off"]: t["read_off"] + t["read_len"]].view(np.int8))
            exp, it = oracle.fill_limited(r, g8, int(t["ref_start"]), int(t["ref_end"]), int(t["min_score"]) - 120, pb, MAXR, MAXC, kind=kind)
            assert out[j].tolist() == exp.tolist(), (tid, j)
            itb += it
        assert ita == itb and np.array_equal(pa, pb), tid
