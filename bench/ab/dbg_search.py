"""debug: search stage of tests/test_mapper_gpu.py::test_map_batch_single_equals_cpu_chain[repeats-14-250], device (all splits) vs oracle"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bbmap_b200 import workloads as wl, search
from bbmap_b200.index import BBIndexCUDA, pack_chromosomes
from bbmap_b200.keyring import KeyRingCUDA, default_cfg
from bbmap_b200.reads import validate_batch
from oracle import oracle as orc
import test_mapper_gpu as T
kind, seed, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
g = T._genome(kind, seed)
cb, co, table = pack_chromosomes([g])
R = wl.make_mapping_reads(cb, co, table, 1500, L=L, seed=seed + 1, sub_rate=0.015, indel_rate=0.02 / 3)
o = orc.get()
eidx = o.index_build(cb, co, 13, -1)
cfg = default_cfg()
es = o.seed_batch(R["bases"], R["qual"], R["off"], cfg, 32)
exp = o.search_batch(eidx, cb, co, R["bases"], es["baseScores"], R["off"], es, quit_after_two_perfects=True)
idx = BBIndexCUDA(cb, co, keylen=13)
for split in (3, 2, 0):
    h, t = search.search_batch(idx, R["bases"], es["baseScores"], R["off"], es, max_sites=search.MAX_SITES, quit_after_two_perfects=True, split=split)
    bad = []
    for i in range(len(h)):
        ns = exp["nsites"][i]
        if h["nsites"][i] != ns or t[i, :ns].tobytes() != exp["sites"][i, :ns].tobytes() or h["best_scores"][i].tolist() != exp["best_scores"][i].tolist() or h["status"][i] != exp["status"][i]:
            bad.append(i)
    print("split", split, "bad reads:", bad[:10], len(bad))
    for i in bad[:3]:
        print(" nkeys", es["nkeys"][i], "dev nsites", h["nsites"][i], "exp", exp["nsites"][i], "best dev", h["best_scores"][i], "exp", exp["best_scores"][i], "status", h["status"][i], exp["status"][i])
        print("  dev", t[i, :h["nsites"][i]][["chrom", "strand", "start", "stop", "hits", "score", "perfect", "semiperfect", "ngaps"]])
        print("  exp", exp["sites"][i, :exp["nsites"][i]][["chrom", "strand", "start", "stop", "hits", "score", "perfect", "semiperfect", "ngaps"]])
