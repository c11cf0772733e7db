"""Known-answer vectors for MultiStateAligner11ts.

Sequences come from the LCG  x=(x*1103515245+12345)&0x7fffffff; base="ACGT"[(x>>16)&3].
The `result`/`iterations` columns were produced by the reference's own C
(jni/MultiStateAligner11tsJNI.c, gcc -O3) — they are reference outputs.  The `score2` and
`match` columns were produced by a restatement of MultiStateAligner11tsJNI.java:376-495,537-658
run on the matrix the reference C filled (no JVM available) — parity unpinned against Java.
"""
import numpy as np


def lcg_seq(n, seed):
    x = seed
    out = []
    for _ in range(n):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        out.append("ACGT"[(x >> 16) & 3])
    return "".join(out)


def B(s):
    return np.frombuffer(s.encode(), dtype=np.int8).copy()


REF = lcg_seq(400, 12345)
MINSCORE = int(0.56 * 9970) - 120  # 5463


def _sub(s, pos, ch=None):
    s = list(s)
    if ch is None:
        ch = "A" if s[pos] != "A" else "C"
    s[pos] = ch
    return "".join(s)


R1 = REF[100:200]
# name, read, a, b, fn, bandwidth, result, iterations, score2, match(RLE)
KATS = [
    ("kat1", R1, 96, 203, "limited", 0, [100, 104, 0, 9970, 0], 5498, [9970, 100, 199, 100, 104, 0], "100m"),
    ("kat1u", R1, 96, 203, "unlimited", 0, [100, 104, 0, 9970], 10800, [9970, 100, 199, 100, 104, 0], "100m"),
    ("kat2", _sub(R1, 50), 96, 203, "limited", 0, [100, 104, 0, 9713, 0], 5409, [9713, 100, 199, 100, 104, 0], "50m1S49m"),
    ("kat3", REF[100:150] + REF[153:203], 96, 206, "limited", 0, [100, 107, 0, 9402, 0], 5518, [9402, 100, 202, 100, 107, 0], "49m3D51m"),
    ("kat3u", REF[100:150] + REF[153:203], 96, 206, "unlimited", 0, [100, 107, 0, 9402], 11100, [9402, 100, 202, 100, 107, 0], "49m3D51m"),
    ("kat4", REF[100:150] + "GT" + REF[150:198], 96, 201, "limited", 0, [100, 102, 0, 9306, 0], 5082, [9306, 100, 197, 100, 102, 0], "50m2I48m"),
    ("kat5", lcg_seq(100, 999), 96, 203, "limited", 0, [100, 1, 0, -2143387648, 1], 1588, None, None),
    ("kat6", REF[100:150] + REF[153:203], 96, 206, "limited", 12, [100, 107, 0, 9402, 0], 2727, [9402, 100, 202, 100, 107, 0], "49m3D51m"),
    ("kat7", _sub(R1, 10, "N"), 96, 203, "limited", 0, [100, 104, 0, 9840, 0], 5372, [9840, 100, 199, 100, 104, 0], "10m1N89m"),
    ("kat8", R1, 100, 199, "limited", 0, [100, 100, 0, 9970, 0], None, [9970, 100, 199, 100, 100, 0], "100m"),
    ("kat9", R1, 104, 203, "unlimited", 0, [100, 96, 0, 9058], None, [9058, 100, 199, 100, 96, 0, 4, 0], "4X96m"),
]


def rle(match):
    s = bytes(np.asarray(match, np.uint8)).decode()
    out = []
    i = 0
    while i < len(s):
        j = i
        while j < len(s) and s[j] == s[i]:
            j += 1
        out.append("%d%s" % (j - i, s[i]))
        i = j
    return "".join(out)
