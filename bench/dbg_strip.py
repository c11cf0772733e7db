import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from bbmap_b200 import workloads as wl
from bbmap_b200.msa import MultiStateAligner11tsCUDA
from oracle import oracle as orc
o=orc.get()
genome = wl.random_genome(50000, seed=21)
reads, tasks = wl.make_msa_tasks(genome, 3000, seed=31, flags=wl.TF_SCORE|wl.TF_TRACEBACK, tight=True, ratio=0.56)
moff = wl.match_offsets(tasks)
exp, emb, cells = o.run_batch(reads, genome, tasks, match_off=moff, threads=8)
for dbg in (0,1,2,3):
  for narrow in (1,0):
    msa=MultiStateAligner11tsCUDA(); msa.set_option("strip_debug",dbg); msa.set_option("narrow",narrow)
    d_ref=msa.load_reference(genome)
    got,gmb=msa.align_batch(reads,d_ref,tasks,match_off=moff)
    bad=[i for i in range(len(tasks)) if got[i].tobytes()!=exp[i].tobytes()]
    print("debug",dbg,"narrow",narrow,"bad",len(bad),"strip_tasks",msa.stat("strip_tasks"), bad[:5])
    if bad and dbg==0 and narrow==0:
        for i in bad[:3]: print(tasks[i], got[i], exp[i], sep="\n")
    msa.close()
