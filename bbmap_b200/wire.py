"""Host-side mirror of the reference's index persistence over the C ABI (csrc/wire.cpp): the files BBMap keeps under
ref/index/<build>/ and ref/genome/<build>/ — Block.write/read (align2/Block.java:74-160), ChromosomeArray.read (dna/ChromosomeArray.java:63-71),
summary.txt (dna/FastaToChromArrays2.java:229-250) — as Java object streams.  Pure host code: works without a device."""
import ctypes as C
import os

import numpy as np

from . import lib as _lib

SUMMARY_DTYPE = np.dtype([("chroms", "<i8"), ("bases", "<i8"), ("defined", "<i8"), ("undefined", "<i8"), ("contigs", "<i8"), ("scaffolds", "<i8"),
                          ("interpad", "<i8"), ("version", "<i4"), ("pad_", "<i4"), ("name", "S256")], align=True)
assert SUMMARY_DTYPE.itemsize == 320


def _check(rc, what):
    if rc != 0:
        raise _lib.BbmError("%s failed (%d): %s" % (what, rc, _lib.load().bbm_wire_last_error().decode()))


def _take(L, ptr, n, dtype):
    """Copy a malloc'ed array out of the library and release it."""
    if n == 0:
        out = np.zeros(0, dtype)
    else:
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()
    L.bbm_wire_free(ptr)
    return out


def block_fname(root_index, min_chrom, max_chrom, k, chrombits, build):
    """IndexMaker4.fname (IndexMaker4.java:477-488)."""
    L = _lib.load()
    buf = C.create_string_buffer(4096)
    _check(L.bbm_wire_block_fname(buf, 4096, os.fsencode(root_index), min_chrom, max_chrom, k, chrombits, build), "bbm_wire_block_fname")
    return buf.value.decode()


def write_int_array(path, a):
    a = np.ascontiguousarray(a, np.int32)
    _check(_lib.load().bbm_wire_write_int_array(os.fsencode(path), a.ctypes.data_as(C.c_void_p), a.size), "bbm_wire_write_int_array")


def read_int_array(path):
    L = _lib.load()
    p = C.c_void_p(); n = C.c_int64(0)
    _check(L.bbm_wire_read_int_array(os.fsencode(path), C.byref(p), C.byref(n)), "bbm_wire_read_int_array")
    return _take(L, p, n.value, np.int32)


def write_block(fname, sites, starts):
    os.makedirs(os.path.dirname(fname) or ".", exist_ok=True)
    sites = np.ascontiguousarray(sites, np.int32); starts = np.ascontiguousarray(starts, np.int32)
    _check(_lib.load().bbm_wire_write_block(os.fsencode(fname), sites.ctypes.data_as(C.c_void_p), sites.size, starts.ctypes.data_as(C.c_void_p), starts.size),
           "bbm_wire_write_block")


def read_block(fname):
    """-> (sites, starts) as Block.read returns them (starts prefix-summed)."""
    L = _lib.load()
    ps = C.c_void_p(); ns = C.c_int64(0); pt = C.c_void_p(); nt = C.c_int64(0)
    _check(L.bbm_wire_read_block(os.fsencode(fname), C.byref(ps), C.byref(ns), C.byref(pt), C.byref(nt)), "bbm_wire_read_block")
    return _take(L, ps, ns.value, np.int32), _take(L, pt, nt.value, np.int32)


def write_chrom(path, chromosome, array, min_index=0, max_index=None, strand=0):
    """ChromosomeArray as FastaToChromArrays2 writes it: array.length == maxIndex+1, minIndex 0, strand Gene.PLUS (0)."""
    a = np.ascontiguousarray(array).view(np.int8)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    _check(_lib.load().bbm_wire_write_chrom(os.fsencode(path), chromosome, a.ctypes.data_as(C.c_void_p), a.size, min_index, a.size - 1 if max_index is None else max_index, strand),
           "bbm_wire_write_chrom")


def read_chrom(path):
    """-> dict(chromosome, array uint8[], minIndex, maxIndex, strand)."""
    L = _lib.load()
    ch = C.c_int32(0); p = C.c_void_p(); n = C.c_int32(0); mn = C.c_int32(0); mx = C.c_int32(0); st = C.c_int8(0)
    _check(L.bbm_wire_read_chrom(os.fsencode(path), C.byref(ch), C.byref(p), C.byref(n), C.byref(mn), C.byref(mx), C.byref(st)), "bbm_wire_read_chrom")
    return {"chromosome": ch.value, "array": _take(L, p, n.value, np.uint8), "minIndex": mn.value, "maxIndex": mx.value, "strand": st.value}


def write_summary(path, **kw):
    g = np.zeros(1, SUMMARY_DTYPE)
    for k, v in kw.items():
        g[k] = v.encode() if isinstance(v, str) else v
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    _check(_lib.load().bbm_wire_write_summary(os.fsencode(path), g.ctypes.data_as(C.c_void_p)), "bbm_wire_write_summary")


def read_summary(path):
    g = np.zeros(1, SUMMARY_DTYPE)
    _check(_lib.load().bbm_wire_read_summary(os.fsencode(path), g.ctypes.data_as(C.c_void_p)), "bbm_wire_read_summary")
    return {k: (g[0][k].decode() if k == "name" else int(g[0][k])) for k in SUMMARY_DTYPE.names if k != "pad_"}


def write_genome(root_genome, build, chrom_bytes, chrom_off, table, name=""):
    """ref/genome/<build>/: chrN.chrom.gz for every chromosome array + summary.txt (FastaToChromArrays2.java:347-353, 229-250)."""
    d = os.path.join(root_genome, str(build))
    defined = 0
    for c in range(len(chrom_off) - 1):
        a = chrom_bytes[chrom_off[c]:chrom_off[c + 1]]
        defined += int(np.isin(a, np.frombuffer(b"ACGTacgt", np.uint8)).sum())
        write_chrom(os.path.join(d, "chr%d.chrom.gz" % (c + 1)), c + 1, a)
    write_summary(os.path.join(d, "summary.txt"), chroms=len(chrom_off) - 1, bases=int(chrom_off[-1]), defined=defined, undefined=int(chrom_off[-1]) - defined,
                  contigs=len(table), scaffolds=len(table), interpad=300, version=5, name=name)


def read_genome(root_genome, build):
    """-> (chrom_bytes uint8[], chrom_off int64[], summary dict) from ref/genome/<build>/ (Data.setGenome + Data.getChromosome)."""
    d = os.path.join(root_genome, str(build))
    s = read_summary(os.path.join(d, "summary.txt"))
    arrs = []
    for c in range(1, s["chroms"] + 1):
        ca = read_chrom(os.path.join(d, "chr%d.chrom.gz" % c))
        arrs.append(ca["array"][: ca["maxIndex"] + 1])
    off = np.zeros(len(arrs) + 1, np.int64); np.cumsum([len(a) for a in arrs], out=off[1:])
    return np.concatenate(arrs), off, s
