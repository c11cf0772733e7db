"""A/B helper: the long-read sub-measurement of bench.py alone (1-kbp reads cut at 500), with per-stage device times and optional context options.
   python bench/ab/long_reads.py [--reads 20000] [--opt key=value ...] [--resident]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bbmap_b200 import workloads as wl
from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
from bbmap_b200.reads import break_reads

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=20000)
ap.add_argument("--genome", type=int, default=4_600_000)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
m = BBMapCUDA([wl.random_genome(a.genome, seed=1)])
for kv in a.opt:
    k, v = kv.split("="); m.L.bbm_set_option(m.h, k.encode(), int(v))
RL = wl.make_long_reads(m.cb, m.co, m.table, a.reads, L=1000, seed=8)
P = break_reads(RL["bases"], RL["qual"], RL["off"], RL["names"], RL["name_off"], 500, 0)
cfg = mapper_cfg(paired=False, sam_text=False)
for _ in range(2):
    d = m.map_batch(P["bases"], P["quality"], P["read_off"], cfg=cfg, match_stride=0)
t0 = time.perf_counter()
for _ in range(a.reps):
    d = m.map_batch(P["bases"], P["quality"], P["read_off"], cfg=cfg, match_stride=0)
dt = (time.perf_counter() - t0) / a.reps
s = d["stats"]
print("pieces/s %.0f  ms/batch %.1f  stages: search %.1f lists %.1f slow %.1f rescue %.1f genmatch %.1f sam %.1f total %.1f  slow_alignments %d realign %d" % (
    len(P["src"]) / dt, dt * 1e3, s["ms_seed_search"], s["ms_lists"], s["ms_slow"], s["ms_rescue"], s["ms_genmatch"], s["ms_sam"], s["ms_total"],
    s["slow_alignments"], s["realign_fills"]))
m.close()
