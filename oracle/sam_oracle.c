/*
 * sam_oracle.c — TEST INFRASTRUCTURE ONLY.
 * C restatement of the SAM record fields BBMap derives from a mapped read, Java-only in the reference:
 *   SamLine(Read, int)                 current/stream/SamLine.java:82-330   (scaffold-relative coordinates, POS, PNEXT, TLEN, RNEXT)
 *   toCigar13 / toCigar14              :600-750     makeFlag  :2134-2151     toMapq  :1709-1723
 *   countLeadingClip / countTrailingClip / countLeadingIndels / countTrailingIndels   :924-1020
 *   Data.isSingleScaffold / scaffoldIndex / scaffoldRelativeLoc     current/dna/Data.java:1089-1140
 *   Read.containsNonM / containsNonNMS   current/stream/Read.java:1815-1863
 * A read is described by the fields SamLine reads from Read: mapped/strand/perfect/ambiguous/secondary/discarded/paired/pairnum flags,
 * chrom/start/stop, mapScore, length and the long-format match string.  PARITY UNPINNED against Java (no JVM).
 */
#pragma GCC optimize ("fp-contract=off")
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "sam_oracle.h"

enum { RF_MAPPED = 1, RF_MINUS = 2, RF_PERFECT = 4, RF_AMBIG = 8, RF_SECONDARY = 16, RF_DISCARDED = 32, RF_PAIRED = 64, RF_PAIRNUM1 = 128 };

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* Arrays.binarySearch: index of key, or -(insertion point)-1 */
static int bsearch_java(const int32_t* a, int n, int key) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) { const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1); if (a[mid] < key) lo = mid + 1; else if (a[mid] > key) hi = mid - 1; else return mid; }
    return -(lo + 1);
}
static int is_single_scaffold(const int32_t* loc, int n, int pad, int loc1, int loc2) {
    if (n < 2) return 1;
    const int idx = bsearch_java(loc, n, loc1 + pad);
    const int scaf = idx >= 0 ? idx : imax(0, (-1 - idx) - 1);
    if (scaf == n - 1) return 1;
    const int lowerBound = loc[scaf] - pad, upperBound = loc[scaf + 1];
    if (loc2 < lowerBound || loc1 > upperBound) return 0;
    return loc2 < upperBound;
}
static int scaffold_index(const int32_t* loc, int n, int pad, int l) {
    if (n < 2) return 0;
    l = l + pad / 2;
    const int idx = bsearch_java(loc, n, l);
    if (idx >= 0) return idx;
    return imax(0, (-1 - idx) - 1);
}
static int count_leading_clip(const int8_t* m, int n) {
    if (!m || n < 1 || m[0] != 'C') return 0;
    int clips = 0, current = 0;
    for (int i = 0; i < n; i++) {
        const int8_t b = m[i];
        if (b >= '0' && b <= '9') current = current * 10 + (b - '0');
        else { if (current > 0) clips = clips + current - 1; current = 0; if (b != 'C') break; clips++; }
    }
    if (current > 0) clips = clips + current - 1;
    return clips;
}
static int count_trailing_clip(const int8_t* m, int n) {
    int clips = 0;
    if (!m) return 0;
    for (int i = n - 1; i >= 0; i--) { if (m[i] == 'C') clips++; else break; }
    return clips;
}
static int count_leading_indels(int rloc, const int8_t* m, int n) {
    if (!m || rloc >= 0) return 0;
    int dels = 0, inss = 0;
    for (int i = 0; i < n && rloc < 0; i++) { const int8_t b = m[i]; if (b == 'D') { dels++; rloc++; } else if (b == 'I') inss++; else rloc++; }
    return dels - inss;
}
/* countTrailingIndels returns 0 whenever rloc >= 0 (SamLine.java:999) — i.e. always, for the b1 >= 0 SamLine passes */
static int count_trailing_indels(int rloc) { return rloc >= 0 ? 0 : 0; }

static int to_mapq(int score, int length, int mapped, int ambig, int penalize) {
    if (!mapped || length < 1) return 0;
    if (ambig && penalize) {
        const float max = 3;
        const float adjusted = (score * max) / (100.0f * length);
        return imax(1, (int)floorf(adjusted + 0.5f));
    } else {
        const float score2 = (score - length * 40) * 1.6f;
        const float max = 1.5f * ((float)(log((double)length) * (1 / log(2.0)))) + 36;
        const float adjusted = (score2 * max) / (100.0f * length);
        return imax(4, (int)floorf(adjusted + 0.5f));
    }
}

static int put_int(int8_t* out, int v) { char tmp[16]; int n = 0; if (v == 0) tmp[n++] = '0'; while (v > 0) { tmp[n++] = (char)('0' + v % 10); v /= 10; } for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i]; return n; }

/* toCigar13 / toCigar14; returns length, -1 for null, -2 for an invalid match character */
static int to_cigar(const int8_t* match, int mlen, int readStart, int readStop, int reflen, int v14, int softClip, int intronLimit, int8_t* out) {
    if (!match || readStart == readStop) return -1;
    int count = 0, o = 0; char mode = '=', lastMode = '=';
    int refloc = readStart;
    for (int mpos = 0; mpos < mlen; mpos++) {
        const int8_t m = match[mpos];
        int sfd = 0;
        if (softClip && (refloc < 0 || refloc >= reflen)) { mode = 'S'; if (m != 'I') refloc++; if (m == 'D') sfd = 1; }
        else if (v14) {
            if (m == 'm' || m == 's') { mode = '='; refloc++; }
            else if (m == 'S') { mode = 'X'; refloc++; }
            else if (m == 'I' || m == 'X' || m == 'Y') mode = 'I';
            else if (m == 'D') { mode = 'D'; refloc++; }
            else if (m == 'C') { mode = 'S'; refloc++; }
            else if (m == 'N' || m == 'B') { mode = 'M'; refloc++; }
            else return -2;
        } else {
            if (m == 'm' || m == 's' || m == 'S' || m == 'N' || m == 'B') { mode = 'M'; refloc++; }
            else if (m == 'I' || m == 'X' || m == 'Y') mode = 'I';
            else if (m == 'D') { mode = 'D'; refloc++; }
            else if (m == 'C') { mode = 'S'; refloc++; }
            else return -2;
        }
        if (mode != lastMode) {
            if (count > 0) { o += put_int(out + o, count); out[o++] = (lastMode == 'D' && count > intronLimit) ? 'N' : lastMode; }
            count = 0; lastMode = mode;
        }
        count++;
        if (sfd) count--;
    }
    o += put_int(out + o, count);
    out[o++] = (mode == 'D' && count > intronLimit) ? 'N' : mode;
    return o;
}

typedef struct { int mapped, paired, has_match, idx, a, b, scaflen, pos0, pos1, gscaf; } side_t;

static void resolve_side(const orc_sam_task* t, const int8_t* match_buf, const int32_t* scaf_off, const int32_t* scaf_loc, const int32_t* scaf_len,
                         int pad, side_t* s) {
    memset(s, 0, sizeof(*s));
    s->mapped = (t->flags & RF_MAPPED) != 0; s->paired = (t->flags & RF_PAIRED) != 0; s->has_match = t->match_len > 0; s->idx = -1; s->gscaf = -1;
    if (s->mapped) {
        const int32_t* loc = scaf_loc + scaf_off[t->chrom - 1]; const int n = scaf_off[t->chrom] - scaf_off[t->chrom - 1];
        if (is_single_scaffold(loc, n, pad, t->start, t->stop)) {
            s->idx = scaffold_index(loc, n, pad, (t->start + t->stop) / 2);
            s->gscaf = scaf_off[t->chrom - 1] + s->idx;
            s->scaflen = scaf_len[s->gscaf];
            s->a = t->start - loc[s->idx];
            s->b = s->a - t->start + t->stop;
        } else { s->mapped = 0; s->paired = 0; s->has_match = 0; }      /* multi-scaffold alignment: SamLine.java:136-141 */
    }
}
static void positions(const orc_sam_task* t, const int8_t* match_buf, side_t* s, int scaflenForTrailing) {
    if (s->mapped) {
        const int8_t* m = s->has_match ? match_buf + t->match_off : 0; const int n = s->has_match ? t->match_len : 0;
        const int clip = count_leading_clip(m, n), ci = count_leading_indels(s->a, m, n), tclip = count_trailing_clip(m, n), tci = count_trailing_indels(s->b);
        s->pos0 = (s->a + 1) + clip + ci;
        s->pos1 = (s->b + 1) - tclip - tci;
        (void)scaflenForTrailing;
    } else { s->pos0 = 0; s->pos1 = 0; }
}

void orc_sam_batch(const orc_sam_task* tasks, int64_t n, const int8_t* match_buf, const int32_t* scaf_off, const int32_t* scaf_loc,
                   const int32_t* scaf_len, int32_t nchroms, const orc_sam_cfg* cfg, orc_sam_out* outs, int8_t* cigar_buf, const int64_t* cigar_off) {
    (void)nchroms;
    for (int64_t i = 0; i < n; i++) {
        const orc_sam_task* t1 = &tasks[i]; const orc_sam_task* t2 = t1->mate >= 0 ? &tasks[t1->mate] : 0;
        orc_sam_out* O = &outs[i];
        side_t s1, s2; memset(&s2, 0, sizeof(s2)); s2.idx = -1; s2.gscaf = -1;
        resolve_side(t1, match_buf, scaf_off, scaf_loc, scaf_len, cfg->inter_scaffold_padding, &s1);
        if (t2) {
            resolve_side(t2, match_buf, scaf_off, scaf_loc, scaf_len, cfg->inter_scaffold_padding, &s2);
            if ((t1->flags & RF_MAPPED) && !s1.mapped) s2.paired = 0;          /* r2.setPaired(false) */
            if ((t2->flags & RF_MAPPED) && !s2.mapped) s1.paired = 0;
        }
        const int sameScaf = (t2 && s1.idx > -1 && s1.idx == s2.idx && t1->chrom == t2->chrom);
        const int minus1 = (t1->flags & RF_MINUS) != 0, minus2 = t2 && (t2->flags & RF_MINUS);
        int flag = 0;
        if (t2) {
            flag |= 0x1;
            if (s1.mapped && s1.has_match && (sameScaf && s1.paired && s2.mapped && s2.has_match)) flag |= 0x2;
            if (t1->flags & RF_PAIRNUM1) flag |= 0x80; else flag |= 0x40;
        }
        if (!s1.mapped) flag |= 0x4;
        if (t2 && !s2.mapped) flag |= 0x8;
        if (minus1) flag |= 0x10;
        if (minus2) flag |= 0x20;
        if (t1->flags & RF_SECONDARY) flag |= 0x100;
        if (t1->flags & RF_DISCARDED) flag |= 0x200;
        positions(t1, match_buf, &s1, s1.scaflen);
        if (s1.mapped) { if (s1.pos1 > s1.scaflen) s1.pos1 = s1.scaflen; if (s1.pos0 < 1) s1.pos0 = 1; }
        if (t2) { positions(t2, match_buf, &s2, s1.scaflen); if (s2.mapped && s2.pos0 < 1) s2.pos0 = 1; }   /* `if(pos1_mate>scaflen){pos1=scaflen;}` touches pos1 only */
        if (t2 && s2.mapped && s2.pos1 > s1.scaflen) s1.pos1 = s1.scaflen;
        int pos, pnext, tlen = 0;
        if (!t2) { pos = s1.pos0; pnext = 0; }
        else if (s1.mapped && s2.mapped) { pos = s1.pos0; pnext = s2.pos0; tlen = sameScaf ? 1 + (imax(s1.pos1, s2.pos1) - imin(s1.pos0, s2.pos0)) : 0; }
        else if (s1.mapped) { pos = s1.pos0; pnext = s1.pos0; }
        else if (s2.mapped) { pos = s2.pos0; pnext = s2.pos0; }
        else { pos = s1.pos0; pnext = s2.pos0; }
        if (!(!t2 || t1->start < t2->start || (t1->start == t2->start && !(t1->flags & RF_PAIRNUM1)))) tlen = -tlen;
        O->flag = flag; O->pos = pos; O->pnext = pnext; O->tlen = tlen;
        O->mapq = to_mapq(t1->score, t1->read_len, s1.mapped, (t1->flags & RF_AMBIG) != 0, cfg->penalize_ambig);
        O->scaffold = s1.mapped ? s1.gscaf : ((t2 && s2.mapped) ? s2.gscaf : -1);
        O->rnext = (!t2 || (!s1.mapped && !s2.mapped)) ? -1 : ((s1.mapped && s2.mapped) ? (sameScaf ? -2 : s2.gscaf) : -2);
        /* cigar */
        int8_t* cg = cigar_buf + cigar_off[i];
        O->cigar_len = -1;
        if (s1.mapped && s1.has_match && t1->read_len > 0) {
            const int8_t* m = match_buf + t1->match_off; const int ml = t1->match_len;
            const int inbounds = (s1.a >= 0 && s1.b < s1.scaflen), perfect = (t1->flags & RF_PERFECT) != 0;
            int nonM = 0, nonNMS = 0;
            for (int k = 0; k < ml; k++) { const int8_t b = m[k]; if (b > '9' && b != 'm') nonM = 1; if (b > '9' && b != 'm' && b != 's' && b != 'N' && b != 'S') nonNMS = 1; }
            if (cfg->version14 ? (inbounds && perfect && !nonM) : (inbounds && (perfect || !nonNMS))) {
                int o = put_int(cg, t1->read_len); cg[o++] = cfg->version14 ? '=' : 'M'; O->cigar_len = o;
            } else O->cigar_len = to_cigar(m, ml, s1.a, s1.b, s1.scaflen, cfg->version14, cfg->soft_clip, cfg->intron_limit, cg);
        }
    }
}
