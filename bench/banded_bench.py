"""BandedAligner at G6 scale: >= 1 M (query, ref) tasks of 150-5000 bp, 0-3 % edits, maxEdits in {2, 5, 26}, Dedupe's
band width rule, all four directions — device time of bbm_banded_batch_dev next to the reference's own jni/BandedAlignerJNI.c on all host cores.

    python bench/banded_bench.py [--unique 50000] [--replicate 20] [--cpu-sample 20000]

The task list holds `unique` distinct pairs (generated pair by pair in Python) repeated `replicate` times: the buffers (unique x ~5 kB) stay far larger
than L2, and the work per task is what the kernel does for 1 M independent pairs.  Roofline: band cells x the 8-lane-op floor of one banded edit-
distance cell (two adds, substitution compare+select, three-way min, early-exit bookkeeping) against the integer issue rate measured in the same run.
Prints one JSON object; bench.py embeds it as `banded`."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bbmap_b200 import lib as _lib, workloads as wl  # noqa: E402

OPS_PER_BAND_CELL = 8


def run(device=0, unique=50_000, replicate=20, cpu_sample=20_000, reps=3, cpu=True):
    import torch
    from bbmap_b200.banded import BAND_OUT_DTYPE
    L = _lib.load()
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    q, rf, bt = wl.make_banded_tasks(unique, seed=6)
    tasks = np.tile(bt, replicate)
    h = C.c_void_p(); _lib.check(L.bbm_init(device, C.byref(h)), "bbm_init")
    pad = lambda a, extra=64: torch.from_numpy(np.concatenate([a, np.zeros(extra, a.dtype)])).to(dev)
    d_q = pad(q); d_r = pad(rf); d_bt = torch.from_numpy(tasks.view(np.uint8)).to(dev)
    d_bo = torch.zeros(len(tasks) * BAND_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    ms = C.c_float(0); times = []
    for it in range(reps + 1):
        _lib.check(L.bbm_banded_batch_dev(h, p(d_q), p(d_r), p(d_bt), p(d_bo), len(tasks), None, C.byref(ms)), "bbm_banded_batch_dev")
        if it:
            times.append(ms.value)
    t = float(np.median(times))
    bo = np.frombuffer(d_bo.cpu().numpy().tobytes(), BAND_OUT_DTYPE)
    width = np.minimum(tasks["max_width"], 2 * tasks["max_edits"] + 1).astype(np.int64)
    rows_done = (bo["rv"][:, 2].astype(np.int64) + 1)                     # lastRow + 1: rows the reference evaluates before its early exit
    cells = int((np.clip(rows_done, 1, None) * width).sum())
    g = C.c_double(0); _lib.check(L.bbm_int_peak(h, 6, C.byref(g)), "bbm_int_peak")
    peak = 2 * g.value
    alg = int((tasks["query_len"].astype(np.int64) + tasks["ref_len"]).sum()) + len(tasks) * (48 + 32)
    out = {"tasks": int(len(tasks)), "distinct_pairs": int(unique), "ms": t, "pairs_per_s": len(tasks) / (t / 1e3), "band_cells": cells, "band_gcups": cells / (t / 1e3) / 1e9,
           "roofline": {"bound": "int-issue", "achieved": cells / (t / 1e3) * OPS_PER_BAND_CELL / 1e9, "peak": peak, "unit": "G lane-ops/s",
                        "frac": cells / (t / 1e3) * OPS_PER_BAND_CELL / 1e9 / max(peak, 1e-9), "ops_per_cell_floor": OPS_PER_BAND_CELL},
           "alg_bytes": alg, "GBps": alg / (t / 1e3) / 1e9, "mean_edits": float(bo["edits"].mean()), "status_nonzero": int((bo["status"] != 0).sum()),
           "workload": "G6: query/ref lengths uniform 150-5000, 0-3 %% edits, maxEdits {2,5,26}, maxWidth = max(min(9, 2*maxEdits+1), 3)|1, four directions, exact 0/1; "
                       "%d tasks = %d distinct pairs x %d" % (len(tasks), unique, replicate)}
    L.bbm_destroy(h)
    if cpu:
        from oracle import oracle as orc
        o = orc.get()
        n = min(cpu_sample, unique)
        threads = os.cpu_count() or 1
        kind = "reference" if o.has_reference else "port"
        t0 = time.perf_counter()
        ro = o.banded_batch(q.view(np.int8), rf.view(np.int8), bt[:n], kind=kind, threads=threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": "first %d distinct pairs, jni/BandedAlignerJNI.c on %d threads, %.2f s" % (n, threads, dt),
                               "identical_to_device": bool(ro[["edits", "rv"]].tobytes() == bo[:n][["edits", "rv"]].tobytes())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--unique", type=int, default=50_000)
    ap.add_argument("--replicate", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=20_000)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(unique=a.unique, replicate=a.replicate, cpu_sample=a.cpu_sample, cpu=not a.no_cpu)))


if __name__ == "__main__":
    main()
