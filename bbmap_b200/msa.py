"""Host-side mirror of the reference's aligner plug-in for the CUDA path.

`MultiStateAligner11tsCUDA` plays the role of current/align2/MultiStateAligner11tsJNI.java behind
MSA.makeMSA (current/align2/MSA.java:38-49): same method names and argument meaning
(fillLimited / fillUnlimited / score / traceback / fillAndScoreLimited), but every method takes a
*batch* of alignments, because a GPU behind a one-call-at-a-time boundary is latency-bound by
construction.  One object per device, like one MSA per mapping thread (AbstractMapThread.java:133-136).
All arithmetic happens in libbbmapcuda.so; this file only marshals buffers.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .workloads import GAPPED_TASK_DTYPE, NOINDEL_TASK_DTYPE, TASK_DTYPE, OUT_DTYPE, TF_RAW_LIMITED, TF_RAW_UNLIMITED, TF_CLAMP, TF_SCORE, TF_TRACEBACK, match_offsets


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class MultiStateAligner11tsCUDA:
    """maxRows/maxColumns mirror the MSA constructor (MSA.java:65-70); bandwidth/bandwidthRatio mirror the statics
    MSA.bandwidth / MSA.bandwidthRatio (MSA.java:864-865)."""

    def __init__(self, maxRows=601, maxColumns=3000, device=0, bandwidth=0, bandwidthRatio=0.0):
        self.L = _lib.load()
        if self.L.bbm_device_count() <= 0:
            raise _lib.BbmError("no CUDA device visible: MultiStateAligner11tsCUDA has no CPU fallback")
        h = C.c_void_p()
        _lib.check(self.L.bbm_init(device, C.byref(h)), "bbm_init")
        self.h = h
        self.maxRows, self.maxColumns = maxRows, maxColumns
        self.iterationsLimited = 0
        self.iterationsUnlimited = 0
        self.set_band(bandwidth, bandwidthRatio)
        self._refs = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.bbm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_band(self, bandwidth, bandwidthRatio):
        self.bandwidth, self.bandwidthRatio = int(bandwidth), float(bandwidthRatio)
        _lib.check(self.L.bbm_set_band(self.h, self.bandwidth, self.bandwidthRatio), "bbm_set_band")

    # -- reference residency (the reference keeps chromosome arrays in memory: dna/Data.java) --
    def load_reference(self, ref_bytes):
        """Upload the concatenated reference arrays once; returns an opaque device pointer."""
        ref = np.ascontiguousarray(ref_bytes).view(np.int8)
        d = C.c_void_p()
        _lib.check(self.L.bbm_upload(self.h, _p(ref), ref.size, C.byref(d)), "bbm_upload")
        return d

    def upload(self, arr):
        a = np.ascontiguousarray(arr)
        d = C.c_void_p()
        _lib.check(self.L.bbm_upload(self.h, _p(a), a.nbytes, C.byref(d)), "bbm_upload")
        return d

    def free(self, d):
        _lib.check(self.L.bbm_free_dev(self.h, d), "bbm_free_dev")

    def set_option(self, key, value):
        _lib.check(self.L.bbm_set_option(self.h, key.encode(), int(value)), "bbm_set_option")

    def stat(self, key):
        return int(self.L.bbm_get_stat(self.h, key.encode()))

    def int_peak(self, kind):
        """Measured giga lane-ops/s of one integer instruction kind (see bbm_int_peak)."""
        g = C.c_double(0)
        _lib.check(self.L.bbm_int_peak(self.h, int(kind), C.byref(g)), "bbm_int_peak")
        return g.value

    @property
    def launches(self):
        return int(self.L.bbm_launch_count(self.h))

    # -- the batched plug-in call (host buffers in, host buffers out) --
    def align_batch(self, reads, d_ref, tasks, match_off=None, outs=None, mbuf=None, account=True):
        """Runs every task (see include/bbmap_cuda.h: bbm_msa_task) and returns (outs, match_buf).
        tasks['flags'] selects fillLimited (Java rule), raw fillLimitedX or raw fillUnlimited, and whether
        score2 / traceback2 follow."""
        reads = np.ascontiguousarray(reads).view(np.int8)
        tasks = np.ascontiguousarray(tasks, TASK_DTYPE)
        if outs is None:
            outs = np.zeros(len(tasks), OUT_DTYPE)
        want_tb = bool(len(tasks)) and bool((tasks["flags"] & TF_TRACEBACK).any())
        if want_tb and match_off is None:
            match_off = match_offsets(tasks)
        if mbuf is None:
            mbuf = np.zeros(int(match_off[-1]) if want_tb else 1, np.int8)
        moff = np.ascontiguousarray(match_off, np.int64) if want_tb else None
        _lib.check(self.L.bbm_msa_batch_host(self.h, _p(reads), reads.size, d_ref, _p(tasks), _p(outs), len(tasks),
                                            _p(mbuf) if want_tb else None, _p(moff) if want_tb else None), "bbm_msa_batch_host")
        if account:           # the reference's iterationsLimited / iterationsUnlimited counters (MSA.java:866-867)
            lim = outs["path"] == 0
            self.iterationsLimited += int(outs["iterations"][lim].sum())
            self.iterationsUnlimited += int(outs["iterations"][~lim].sum())
        return outs, mbuf

    def align_batch_gapped(self, reads, d_ref, gtasks, gaps, match_off):
        """MSA.fillAndScoreLimited(read, ref, start-thresh, stop+thresh, minScore, gaps) [+ traceback] for a batch
        (MSA.java:103-134): tasks with ngaps>0 are aligned against the gapped reference makeGref builds
        (MultiStateAligner11tsJNI.java:668-757), on the device; score[1:3] come back in chromosome coordinates."""
        reads = np.ascontiguousarray(reads).view(np.int8)
        gtasks = np.ascontiguousarray(gtasks, GAPPED_TASK_DTYPE)
        gaps = np.ascontiguousarray(gaps, np.int32)
        outs = np.zeros(len(gtasks), OUT_DTYPE)
        moff = np.ascontiguousarray(match_off, np.int64)
        mbuf = np.zeros(max(int(moff[-1]), 1), np.int8)
        _lib.check(self.L.bbm_msa_gapped_batch_host(self.h, _p(reads), reads.size, d_ref, _p(gtasks), _p(gaps) if len(gaps) else None,
                                                   len(gaps), _p(outs), len(gtasks), _p(mbuf), _p(moff)), "bbm_msa_gapped_batch_host")
        return outs, mbuf

    def align_batch_dev(self, d_reads, d_ref, d_tasks, d_outs, ntasks, d_match, d_moff, max_rows, max_cols, stream=None):
        """Everything resident on the device (device pointers as ints / c_void_p). Returns device ms (CUDA events)."""
        ms = C.c_float(0)
        _lib.check(self.L.bbm_msa_batch_dev(self.h, d_reads, d_ref, d_tasks, d_outs, ntasks, d_match, d_moff,
                                           max_rows, max_cols, stream, C.byref(ms)), "bbm_msa_batch_dev")
        return ms.value

    def scoreNoIndels(self, reads, d_ref, tasks, match_off=None):
        """MSA.scoreNoIndels for a batch of (read, site) pairs (…JNI.java:1033-1089); with match_off also the match strings
        of scoreNoIndelsAndMakeMatchString (:1243-1318).  Returns (scores, match_buf)."""
        reads = np.ascontiguousarray(reads).view(np.int8)
        tasks = np.ascontiguousarray(tasks, NOINDEL_TASK_DTYPE)
        scores = np.zeros(len(tasks), np.int32)
        mbuf = np.zeros(int(match_off[-1]) if match_off is not None else 1, np.int8)
        moff = None if match_off is None else np.ascontiguousarray(match_off, np.int64)
        _lib.check(self.L.bbm_noindel_batch_host(self.h, _p(reads), reads.size, d_ref, _p(tasks), _p(scores),
                                                _p(mbuf) if moff is not None else None, _p(moff) if moff is not None else None, len(tasks)),
                   "bbm_noindel_batch_host")
        return scores, mbuf

    # -- reference-named conveniences over align_batch (argument meaning as in MSA.java / …JNI.java) --
    def fillAndScoreLimited(self, reads, d_ref, tasks):
        """MSA.fillAndScoreLimited (MSA.java:103-134) for a batch: clamp, fillLimited, score. score_len==0 ⇔ null."""
        t = tasks.copy()
        t["flags"] = (t["flags"] & ~(TF_RAW_LIMITED | TF_RAW_UNLIMITED | TF_TRACEBACK)) | TF_CLAMP | TF_SCORE
        return self.align_batch(reads, d_ref, t)[0]

    def fillLimited(self, reads, d_ref, tasks):
        t = tasks.copy()
        t["flags"] = t["flags"] & ~(TF_RAW_LIMITED | TF_RAW_UNLIMITED | TF_SCORE | TF_TRACEBACK)
        return self.align_batch(reads, d_ref, t)[0]

    def fillUnlimited(self, reads, d_ref, tasks):
        t = tasks.copy()
        t["flags"] = (t["flags"] & ~(TF_RAW_LIMITED | TF_SCORE | TF_TRACEBACK)) | TF_RAW_UNLIMITED
        return self.align_batch(reads, d_ref, t)[0]
