"""TEST INFRASTRUCTURE ONLY — a second, independent statement of MultiStateAligner11tsJNI.score2 / traceback2 in plain Python.

Written from the Java text (current/align2/MultiStateAligner11tsJNI.java:376-495 traceback2, :537-658 score2), not from
oracle/msa_oracle.c and not from the kernels: it walks the flat `packed[3][maxRows+1][maxColumns+1]` matrix that the
REFERENCE'S OWN C (oracle/_ref/libbbref.so, compiled unmodified from jni/MultiStateAligner11tsJNI.c) has just filled, the way
the Java methods do after the JNI call returns.  tests/test_walk_independent.py compares it with the C restatement (and through
it with the CUDA predecessor-code walk) on >= 10^4 random alignments, so a slip in reading the Java made once in the C
restatement does not pass unnoticed."""

TIMEMASK = (1 << 11) - 1
SCOREMASK = ~TIMEMASK
SCOREOFFSET = 11
MODE_MS, MODE_DEL, MODE_INS = 0, 1, 2
GAPC = ord('-')
GAPLEN = 128
POINTSoff_NOREF = 0            # MultiStateAligner11tsJNI.java:1497 (POINTS_NOREF=0)

_DEFINED = frozenset(b"ACGTUacgtu")      # AminoAcid.isFullyDefined: baseToNumber[b]>=0 (dna/AminoAcid.java:365-367,615-624)


class Matrix:
    """packed[state][row][col] over the flat int32 array the C wrote."""

    def __init__(self, packed, maxRows, maxColumns):
        self.p = packed
        self.plane = (maxRows + 1) * (maxColumns + 1)
        self.stride = maxColumns + 1

    def at(self, state, row, col):
        return int(self.p[state * self.plane + row * self.stride + col])

    def score(self, state, row, col):
        return self.at(state, row, col) & SCOREMASK

    def time(self, state, row, col):
        return self.at(state, row, col) & TIMEMASK


def _previous_state(M, state, row, col):
    """The predecessor choice shared by score2 (:573-611) and traceback2 (:391-447): stay while the streak counter says so,
    else compare the raw neighbour scores, diagonal first."""
    if M.time(state, row, col) > 1:
        return state
    if state == MODE_MS:
        d = M.score(MODE_MS, row - 1, col - 1); e = M.score(MODE_DEL, row - 1, col - 1); i = M.score(MODE_INS, row - 1, col - 1)
        if d >= e and d >= i:
            return MODE_MS
        return MODE_DEL if e >= i else MODE_INS
    if state == MODE_DEL:
        return MODE_MS if M.score(MODE_MS, row, col - 1) >= M.score(MODE_DEL, row, col - 1) else MODE_DEL
    return MODE_MS if M.score(MODE_MS, row - 1, col) >= M.score(MODE_INS, row - 1, col) else MODE_INS


def score2(M, rows, columns, refStartLoc, refEndLoc, maxRow, maxCol, maxState):
    """-> list of 6 or 8 ints (:537-658)."""
    row, col, state = maxRow, maxCol, maxState
    score = M.score(maxState, maxRow, maxCol)
    if row < rows:
        difR = rows - row; difC = columns - col
        while difR > difC:
            score += POINTSoff_NOREF; difR -= 1
        row += difR; col += difR
    bestRefStop = refStartLoc + col - 1
    stateTime = 0
    while row > 0 and col > 0:
        prev = _previous_state(M, state, row, col)
        if state == MODE_MS:
            row -= 1; col -= 1
        elif state == MODE_DEL:
            col -= 1
        else:
            row -= 1
        if col < 0:
            break
        stateTime = stateTime + 1 if state == prev else 0
        state = prev
    if row > col:
        col -= row
    bestRefStart = refStartLoc + col
    score >>= SCOREOFFSET
    padLeft = padRight = 0
    if bestRefStart < refStartLoc:
        padLeft = max(0, refStartLoc - bestRefStart)
    elif bestRefStart == refStartLoc and state == MODE_INS:
        padLeft = stateTime
    if bestRefStop > refEndLoc:
        padRight = max(0, bestRefStop - refEndLoc)
    elif bestRefStop == refEndLoc and maxState == MODE_INS:
        padRight = M.time(maxState, maxRow, maxCol)
    out = [score, bestRefStart, bestRefStop, maxRow, maxCol, maxState]
    if padLeft > 0 or padRight > 0:
        out += [padLeft, padRight]
    return out


def traceback2(M, read, ref, columns, refStartLoc, row, col, state):
    """-> bytes, the match string (:376-495)."""
    out = bytearray()
    gaps = 0
    while row > 0 and col > 0:
        prev = _previous_state(M, state, row, col)
        if state == MODE_MS:
            c = read[row - 1]; r = ref[refStartLoc + col - 1]
            if c == r:
                out.append(ord('m'))
            elif c not in _DEFINED or r not in _DEFINED:
                out.append(ord('N'))
            else:
                out.append(ord('S'))
            row -= 1; col -= 1
        elif state == MODE_DEL:
            if ref[refStartLoc + col - 1] == GAPC:
                out.append(GAPC); gaps += 1
            else:
                out.append(ord('D'))
            col -= 1
        else:
            out.append(ord('X') if col == 0 else (ord('Y') if col >= columns else ord('I')))
            row -= 1
        state = prev
    if col != row:
        while row > 0:
            out.append(ord('X')); row -= 1; col -= 1
    out.reverse()
    if gaps == 0:
        return bytes(out)
    return bytes(out).replace(b"-", b"D" * GAPLEN)
