"""Regenerates tests/golden/phix.npz from the reference's own shipped test inputs (BASELINE configs[0]):
/root/reference/resources/phix174_ill.ref.fa.gz + sample1.fq.gz + sample2.fq.gz (100 synthetic 100-bp pairs whose names carry
the true origin: <id>_chr<chrom>_<strand>_<start>_<stop>_<scafpos>_<scaffold>, chromosome coordinates with the 8000-N lead pad
of dna/FastaToChromArrays2.java:569).  Data only — no reference source is copied.  Run in the build container (the
reference mount does not exist on the GPU box); the .npz is committed."""
import gzip
import os

import numpy as np

REF = "/root/reference/resources"
HERE = os.path.dirname(os.path.abspath(__file__))


def fasta(path):
    seq = []
    for line in gzip.open(path, "rt"):
        if not line.startswith(">"):
            seq.append(line.strip())
    return np.frombuffer("".join(seq).encode(), np.uint8)


def fastq(path):
    names, bases, quals = [], [], []
    lines = gzip.open(path, "rt").read().split("\n")
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i][1:]); bases.append(lines[i + 1].strip()); quals.append(lines[i + 3].strip())
    return names, bases, quals


def main():
    genome = fasta(os.path.join(REF, "phix174_ill.ref.fa.gz"))
    out = {"genome": genome}
    for tag, f in (("r1", "sample1.fq.gz"), ("r2", "sample2.fq.gz")):
        names, bases, quals = fastq(os.path.join(REF, f))
        truth = np.array([[int(nm.split("_")[1][3:])] + [int(x) for x in nm.split("_")[2:6]] for nm in names], np.int32)
        out[tag + "_bases"] = np.frombuffer("".join(bases).encode(), np.uint8)
        out[tag + "_qual"] = np.frombuffer("".join(quals).encode(), np.uint8) - 33      # phred, ASCII offset removed
        off = np.zeros(len(bases) + 1, np.int64); np.cumsum([len(b) for b in bases], out=off[1:])
        out[tag + "_off"] = off
        out[tag + "_truth"] = truth          # chrom, strand, start, stop, scaffold position
    np.savez_compressed(os.path.join(HERE, "phix.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
