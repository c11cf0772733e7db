import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bbmap_b200 import workloads as wl
from bbmap_b200.mapper import BBMapCUDA, mapper_cfg
from oracle import chain, oracle as orc
import test_mapper_gpu as T
from kat import rle
scafs = [wl.random_genome(150_000, seed=51), wl.random_genome(60_000, seed=52)]
m = BBMapCUDA(scafs, names=["s1", "s2"])
rng = np.random.Generator(np.random.PCG64(53))
bases, qual, off = T._edge_reads(scafs, rng)
o = orc.get(); idx = o.index_build(m.cb, m.co, 13, -1)
ref = chain.map_single(o, idx, m.cb, m.co, m.table, bases, qual, off)
dev = m.map_batch(bases, qual, off, cfg=mapper_cfg(paired=False), match_stride=ref["match_stride"])
ms = ref["match_stride"]
for r in range(len(off) - 1):
    a, b = dev["recs"][r], ref["recs"][r]
    if any(a[f] != b[f] for f in T.REC_FIELDS):
        print("read", r, "len", off[r + 1] - off[r])
        print(" dev", a); print(" ref", b)
        dm = dev["match"][r * ms: r * ms + a["match_len"]]; om = ref["match"][r * ms: r * ms + b["match_len"]]
        print(" dev match", rle(dm)); print(" ref match", rle(om))
