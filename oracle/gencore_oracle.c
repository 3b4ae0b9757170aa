/*
 * gencore_oracle.c — plain-C restatement of OpenGene/gencore v0.17.2's consensus hot path.
 * TEST INFRASTRUCTURE ONLY (see gencore_oracle.h).  Each function cites the reference lines it
 * restates (paths under /root/reference/src).  It follows the reference literally — in-place
 * mutation of a private copy of the payload, lazily computed per-pair scores, the same loop
 * orders — so that it can be read side by side with the C++; it is written for checking, not
 * for speed.
 */
#include "gencore_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- string-level helpers */

/* Cluster::umiDiff, cluster.cpp:41-53 */
int gco_umi_diff(const char *u1, const char *u2) {
    int len1 = (int)strlen(u1), len2 = (int)strlen(u2);
    int diff = abs(len1 - len2);
    int m = len1 < len2 ? len1 : len2;
    for (int i = 0; i < m; i++)
        if (u1[i] != u2[i]) diff++;
    return diff;
}

/* util.h:59-88 split(str, ret, "_") restated: returns number of parts, part k = [beg[k], end[k]) */
static int split_underscore(const char *s, int len, int *beg, int *end, int cap) {
    int n = 0;
    if (len == 0) return 0;
    int pos = 0;
    while (pos < len && s[pos] == '_') pos++; /* find_first_not_of */
    if (pos >= len) return 0;
    for (;;) {
        int comma = -1;
        for (int k = pos; k < len; k++)
            if (s[k] == '_') { comma = k; break; }
        if (comma >= 0) {
            if (n < cap) { beg[n] = pos; end[n] = comma; }
            n++;
            pos = comma + 1; /* may equal len: the next round pushes an empty tail part */
        } else {
            if (n < cap) { beg[n] = pos; end[n] = len; }
            n++;
            break;
        }
    }
    return n;
}

/* Cluster::isDuplex, cluster.cpp:246-258 */
int gco_is_duplex(const char *u1, const char *u2) {
    int b1[3], e1[3], b2[3], e2[3];
    int n1 = split_underscore(u1, (int)strlen(u1), b1, e1, 3);
    int n2 = split_underscore(u2, (int)strlen(u2), b2, e2, 3);
    if (n1 != 2 || n2 != 2) return 0;
    int a0 = e1[0] - b1[0], a1 = e1[1] - b1[1], c0 = e2[0] - b2[0], c1 = e2[1] - b2[1];
    if (a0 == c1 && a1 == c0 && memcmp(u1 + b1[0], u2 + b2[1], a0) == 0 && memcmp(u1 + b1[1], u2 + b2[0], a1) == 0)
        return 1;
    return 0;
}

static int is_umi_char(char c) { return c == 'A' || c == 'T' || c == 'C' || c == 'G' || c == '_'; }

/* BamUtil::getUMI(string qname, const string& prefix), bamutil.cpp:40-112. Returns length or -1. */
int gco_get_umi(const char *qname, const char *prefix, char *out, int cap) {
    int len = (int)strlen(qname), prefixLen = (int)strlen(prefix);
    out[0] = 0;
    if (prefixLen > 0) { /* bamutil.cpp:45-63: find_last_of(prefix) = last char that is ANY char of prefix */
        int pos = -1;
        for (int i = len - 1; i >= 0 && pos < 0; i--)
            if (strchr(prefix, qname[i])) pos = i;
        if (pos < 0) return 0;
        int start = pos + 2, umiLen = 0;
        for (int sep = start; sep < len; sep++) {
            if (!is_umi_char(qname[sep])) break;
            umiLen++;
        }
        if (start > len) return 0; /* substr(start) with start > size would throw; start == size gives "" */
        if (umiLen + 1 > cap) return -1;
        memcpy(out, qname + start, umiLen);
        out[umiLen] = 0;
        return umiLen;
    }
    int sep, foundSep = 0;
    for (sep = len - 1; sep >= 0; sep--)
        if (qname[sep] == ':') { foundSep = 1; break; }
    if (!foundSep || sep + prefixLen >= len - 1) return 0; /* prefixLen == 0 here */
    int start = sep + 1 + prefixLen;
    if (start < len - 1 && qname[start] == '_') start++;
    int underscores = 0;
    for (int i = start; i < len; i++) {
        char c = qname[i];
        if (!is_umi_char(c)) return 0;
        if (c == '_' && ++underscores > 1) return 0;
    }
    if (len - start + 1 > cap) return -1;
    memcpy(out, qname + start, len - start);
    out[len - start] = 0;
    return len - start;
}

static int umi_char_code(char c) {
    switch (c) {
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 3;
        case 'T': return 4;
        case '_': return 5;
        default: return 0;
    }
}
static const char UMI_CODE_CHAR[8] = {0, 'A', 'C', 'G', 'T', '_', '?', '?'};

int gco_encode_umi(const char *umi, uint64_t *words, int n_words) {
    int len = (int)strlen(umi);
    for (int w = 0; w < n_words; w++) words[w] = 0;
    if (len > 16 * n_words) return -1;
    for (int k = 0; k < len; k++) {
        int code = umi_char_code(umi[k]);
        if (!code) return -1;
        words[k >> 4] |= (uint64_t)code << (60 - 4 * (k & 15));
    }
    return len;
}

static void decode_umi(const uint64_t *words, int n_words, char *out) {
    int k = 0;
    for (; k < 16 * n_words; k++) {
        int code = (int)((words[k >> 4] >> (60 - 4 * (k & 15))) & 0xF);
        if (!code) break;
        out[k] = UMI_CODE_CHAR[code & 7];
    }
    out[k] = 0;
}

/* FastaReader::to4bits + base2bits, fastareader.cpp:106-113,139-152 */
void gco_pack_genome(const char *bases, int64_t n, uint8_t *packed4) {
    memset(packed4, 0, (size_t)((n + 1) / 2));
    for (int64_t i = 0; i < n; i++) {
        char b = bases[i];
        uint8_t bits = b == 'A' ? 1 : b == 'T' ? 2 : b == 'C' ? 3 : b == 'G' ? 4 : 0;
        if (i % 2 == 0) packed4[i / 2] |= bits;
        else packed4[i / 2] |= (uint8_t)(bits << 4);
    }
}

/* ---------------------------------------------------------------- working state */

typedef struct {
    const gcb_options *opt;
    const gco_genome *genome;
    const gcb_batch *b;
    uint8_t *work;      /* private, mutable copy of the payload (the reference mutates records in place) */
    signed char **score; /* [2*n_pairs] Pair::mLeftScore / mRightScore, lazily allocated */
    int32_t *have;      /* [2*n_pairs] 1 while Pair::mLeft/mRight is non-NULL (0 once stolen as template) */
} W;

static inline const gcb_read_desc *RD(const W *w, int slot) { return &w->b->reads[slot]; }
static inline uint8_t *QUAL(const W *w, int slot) { return w->work + RD(w, slot)->data_off; }
static inline uint8_t *SEQ(const W *w, int slot) {
    return w->work + RD(w, slot)->data_off + GCB_ALIGN4(RD(w, slot)->l_qseq);
}
static inline const uint32_t *CIG(const W *w, int slot) { return w->b->cigar + RD(w, slot)->cigar_off; }
static inline uint8_t base_at(const uint8_t *seq, int i) { /* bam_get_seq nibble order */
    return (i % 2 == 1) ? (seq[i / 2] & 0xF) : ((seq[i / 2] >> 4) & 0xF);
}

#define OP(c) ((int)((c) & 0xF))
#define OPLEN(c) ((uint32_t)((c) >> 4))
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5

/* bamutil.cpp:290-291 */
static const int QUERY_CONSUM[16] = {1, 1, 0, 0, 1, 0, 0, 1, 1, 0};
static const int REFERENCE_CONSUM[16] = {1, 0, 1, 1, 0, 0, 0, 1, 1, 0};

/* BamUtil::getRefOffset, bamutil.cpp:293-314 */
static int get_ref_offset(const uint32_t *cig, int n, int bampos) {
    int ref = 0, query = 0;
    for (int i = 0; i < n; i++) {
        int op = OP(cig[i]);
        uint32_t len = OPLEN(cig[i]);
        query += len * QUERY_CONSUM[op];
        ref += len * REFERENCE_CONSUM[op];
        if (query > bampos) {
            if (op == BAM_CINS || op == BAM_CSOFT_CLIP) return -1;
            return ref - REFERENCE_CONSUM[op] * (query - bampos);
        }
    }
    return -1;
}

/* BamUtil::getMOffsetAndLen, bamutil.cpp:316-336: first M block only */
static void get_m_offset_and_len(const uint32_t *cig, int n, int *MOffset, int *MLen) {
    int query = 0;
    for (int i = 0; i < n; i++) {
        int op = OP(cig[i]);
        uint32_t len = OPLEN(cig[i]);
        if (op == BAM_CMATCH) { *MOffset = query; *MLen = (int)len; return; }
        query += len * QUERY_CONSUM[op];
    }
    *MOffset = 0;
    *MLen = 0;
}

/* BamUtil::getRightRefPos, bamutil.cpp:379-383 (bam_cigar2rlen = sum of ref-consuming op lengths) */
static int64_t get_right_ref_pos(const W *w, int slot) {
    const gcb_read_desc *r = RD(w, slot);
    if (r->pos < 0) return -1;
    int64_t l = 0;
    const uint32_t *cig = CIG(w, slot);
    for (int k = 0; k < r->n_cigar; k++)
        if (REFERENCE_CONSUM[OP(cig[k])]) l += OPLEN(cig[k]);
    return r->pos + l;
}

/* BamUtil::isPartOf, bamutil.cpp:204-255 */
static int is_part_of(const uint32_t *cp, int np, const uint32_t *cw, int nw, int isLeft) {
    if (nw < np) return 0;
    for (int i = 0; i < np; i++) {
        uint32_t vp = isLeft ? cp[i] : cp[np - i - 1];
        uint32_t vw = isLeft ? cw[i] : cw[nw - i - 1];
        if (OP(vp) != OP(vw)) return 0;
        if (OPLEN(vp) > OPLEN(vw)) return 0;
        if (OPLEN(vp) < OPLEN(vw)) {
            if (i != np - 1) {
                if (i != np - 2) return 0;
                int next = i + 1;
                uint32_t vn = isLeft ? cp[next] : cp[np - next - 1];
                if (OP(vn) != BAM_CHARD_CLIP) return 0;
            }
        }
    }
    return 1;
}
static int is_part_of_slots(const W *w, int part, int whole, int isLeft) {
    return is_part_of(CIG(w, part), RD(w, part)->n_cigar, CIG(w, whole), RD(w, whole)->n_cigar, isLeft);
}

/* Pair::qual2score, pair.cpp:77-86 */
static signed char qual2score(const gcb_options *o, uint8_t q) {
    if (o->high_quality <= q) return (signed char)o->score_high;
    else if (o->moderate_quality <= q) return (signed char)o->score_moderate;
    else if (o->low_quality <= q) return (signed char)o->score_low;
    else return (signed char)o->score_bad;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* Pair::computeScore, pair.cpp:88-172 (incl. the in-place qual rewrite at 158-159) */
static void compute_score(W *w, int pair) {
    const gcb_options *o = w->opt;
    int L = 2 * pair, R = 2 * pair + 1;
    if (w->have[L] && !w->score[L]) {
        int n = RD(w, L)->l_qseq;
        w->score[L] = (signed char *)malloc(n > 0 ? n : 1);
        memset(w->score[L], o->score_moderate, n);
    }
    if (w->have[R] && !w->score[R]) {
        int n = RD(w, R)->l_qseq;
        w->score[R] = (signed char *)malloc(n > 0 ? n : 1);
        memset(w->score[R], o->score_moderate, n);
    }
    if (w->score[L] && w->score[R] && w->have[L] && w->have[R]) {
        /* note: in the reference mLeft/mRight are both still present whenever both score arrays
           are first created together; `have` can only drop after this function has run once. */
        int leftMOffset, leftMLen, rightMOffset, rightMLen;
        get_m_offset_and_len(CIG(w, L), RD(w, L)->n_cigar, &leftMOffset, &leftMLen);
        get_m_offset_and_len(CIG(w, R), RD(w, R)->n_cigar, &rightMOffset, &rightMLen);
        if (leftMLen > 0 && rightMLen > 0) {
            int posDis = RD(w, R)->pos - RD(w, L)->pos;
            int leftStart, rightStart, cmpLen;
            if (posDis >= 0) {
                leftStart = leftMOffset + posDis;
                rightStart = rightMOffset;
                cmpLen = imin(leftMLen - posDis, rightMLen);
            } else {
                leftStart = leftMOffset;
                rightStart = rightMOffset - posDis;
                cmpLen = imin(leftMLen, rightMLen + posDis);
            }
            uint8_t *lseq = SEQ(w, L), *rseq = SEQ(w, R), *lqual = QUAL(w, L), *rqual = QUAL(w, R);
            int ll = RD(w, L)->l_qseq, rl = RD(w, R)->l_qseq;
            for (int i = 0; i < imin(ll, leftStart); i++) w->score[L][i] = qual2score(o, lqual[i]);
            for (int i = imax(0, leftStart + cmpLen); i < ll; i++) w->score[L][i] = qual2score(o, lqual[i]);
            for (int i = 0; i < imin(rl, rightStart); i++) w->score[R][i] = qual2score(o, rqual[i]);
            for (int i = imax(0, rightStart + cmpLen); i < rl; i++) w->score[R][i] = qual2score(o, rqual[i]);
            for (int i = 0; i < cmpLen; i++) {
                int l = leftStart + i, r = rightStart + i;
                if (l < 0 || l >= ll || r < 0 || r >= rl) continue; /* reference would read out of bounds */
                uint8_t lq = lqual[l], rq = rqual[r];
                uint8_t lbase = base_at(lseq, l), rbase = base_at(rseq, r);
                if (lbase == rbase) {
                    uint8_t q = (uint8_t)((lq + rq) / 2);
                    signed char score = (signed char)(qual2score(o, q) + 4);
                    w->score[L][l] = score;
                    w->score[R][r] = score;
                } else {
                    lqual[l] = (uint8_t)imax(0, (int)lq - (int)rq);
                    rqual[r] = (uint8_t)imax(0, (int)rq - (int)lq);
                    if (lq >= rq) {
                        w->score[L][l] = (signed char)(qual2score(o, (uint8_t)(lq - rq)) - 3);
                        w->score[R][r] = 0;
                    } else {
                        w->score[L][l] = 0;
                        w->score[R][r] = (signed char)(qual2score(o, (uint8_t)(rq - lq)) - 3);
                    }
                }
            }
        }
    }
}

/* Pair::getLeftScore / getRightScore, pair.cpp:174-186 */
static signed char *get_score(W *w, int pair, int side) {
    if (!w->score[2 * pair + side]) compute_score(w, pair);
    return w->score[2 * pair + side];
}

/* Reference::getData, reference.cpp:33-71 (the mLast* cache does not change the result) */
static const uint8_t *ref_get_data(const W *w, int contig, int64_t pos, int64_t len, int64_t *contig_len) {
    const gco_genome *g = w->genome;
    if (!g || !g->packed4) return NULL;
    if (contig < 0 || contig >= g->n_contigs) return NULL;
    if (pos + len >= g->contig_len[contig]) return NULL;
    *contig_len = g->contig_len[contig];
    return g->packed4 + g->contig_off[contig];
}

/* FastaReader::getBase + bits2base, fastareader.cpp:115-128 */
static char ref_get_base(const uint8_t *refdata, int64_t refpos) {
    uint8_t two = refdata[refpos / 2];
    uint8_t bits = (refpos % 2 == 0) ? (two & 0x0F) : ((two & 0xF0) >> 4);
    static const char bases[5] = {'N', 'A', 'T', 'C', 'G'};
    return bits >= 5 ? 'N' : bases[bits];
}

/* BamUtil::base2fourbits, bamutil.cpp:167-183 (only A/C/G/T reach it here) */
static uint8_t base2fourbits(char base) {
    switch (base) {
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 4;
        case 'T': return 8;
        default: return 15;
    }
}

/* Group::makeConsensus, group.cpp:320-579. reads[0] == out. Returns diff; *mismatch_inc_out = mismatchInc. */
static int make_consensus(W *w, const int *reads, signed char **scores, int nreads, int out, int isLeft,
                          int contig, int *mismatch_inc_out) {
    const gcb_options *o = w->opt;
    const gcb_read_desc *od = RD(w, out);
    int diff = 0, mismatchInc = 0;
    int seqbytes = (od->l_qseq + 1) >> 1, qualbytes = od->l_qseq;
    uint8_t *seqBak = (uint8_t *)malloc(seqbytes + 1), *qualBak = (uint8_t *)malloc(qualbytes + 1);
    memcpy(seqBak, SEQ(w, out), seqbytes);
    memcpy(qualBak, QUAL(w, out), qualbytes);

    int *lenDiff = (int *)malloc(sizeof(int) * nreads);
    for (int r = 0; r < nreads; r++) { /* group.cpp:339-349 */
        int d = RD(w, reads[r])->l_qseq - od->l_qseq;
        if (d != 0) {
            if (RD(w, reads[r])->pos == od->pos && is_part_of_slots(w, out, reads[r], 1)) d = 0;
        }
        lenDiff[r] = d;
    }
    uint8_t *outdata = SEQ(w, out), *outqual = QUAL(w, out);
    int len = od->l_qseq;
    if (od->n_cigar == 0) { /* group.cpp:354-360 */
        for (int r = 0; r < nreads; r++)
            if (RD(w, reads[r])->l_qseq < len) len = RD(w, reads[r])->l_qseq;
    }
    const uint8_t *refdata = NULL;
    int64_t contigLen = 0;
    if (od->isize != 0) /* group.cpp:362-367 */
        refdata = ref_get_data(w, contig, od->pos, get_ref_offset(CIG(w, out), od->n_cigar, len - 1) + 1, &contigLen);

    for (int i = 0; i < len; i++) { /* group.cpp:369-526 */
        int counts[16] = {0}, baseScores[16] = {0}, quals[16] = {0};
        uint8_t topQuals[16] = {0};
        int totalScore = 0;
        for (int r = 0; r < nreads; r++) {
            int readpos = i;
            if (!isLeft) readpos = i + lenDiff[r];
            if (readpos < 0 || readpos >= RD(w, reads[r])->l_qseq) continue; /* reference: out-of-bounds read */
            uint8_t qual = QUAL(w, reads[r])[readpos];
            uint8_t base = base_at(SEQ(w, reads[r]), readpos);
            counts[base]++;
            baseScores[base] += scores[r][readpos];
            totalScore += scores[r][readpos];
            quals[base] += qual;
            if (qual > topQuals[base]) topQuals[base] = qual;
        }
        uint8_t topBase = 0;
        int topScore = -0x7FFFFFFF;
        for (uint8_t b = 0; b < 16; b++) {
            if (baseScores[b] > topScore || (baseScores[b] == topScore && quals[b] >= quals[topBase])) {
                topScore = baseScores[b];
                topBase = b;
            }
        }
        int topNum = counts[topBase];
        uint8_t topQual = topQuals[topBase];
        uint8_t secBase = 0;
        int secScore = -0x7FFFFFFF;
        for (uint8_t b = 0; b < 16; b++) {
            if (b == topBase) continue;
            if (baseScores[b] > secScore || (baseScores[b] == secScore && quals[b] >= quals[secBase])) {
                secScore = baseScores[b];
                secBase = b;
            }
        }
        int secNum = counts[secBase];
        int needToCheckRef = 0;
        if (secNum == 0) { /* group.cpp:421-428: fast path writes the qual only, NOT the base */
            if (topScore >= o->base_score_req && topQual >= o->moderate_quality) {
                outqual[i] = topQual;
                continue;
            } else
                needToCheckRef = 1;
        }
        char refbase = 0;
        if (refdata) {
            int refpos = get_ref_offset(CIG(w, out), od->n_cigar, i);
            if (refpos >= 0) refbase = ref_get_base(refdata, (int64_t)od->pos + refpos);
        }
        if (refbase != 'A' && refbase != 'T' && refbase != 'C' && refbase != 'G') refbase = 0;
        if (secNum == 1) { /* group.cpp:442-457 */
            if (quals[secBase] <= o->low_quality) {
                if (topNum < 2 && topQual < o->high_quality) needToCheckRef = 1;
            } else {
                if (topNum < 3 || topQual < o->high_quality) needToCheckRef = 1;
            }
        }
        if (secNum > 1) { /* group.cpp:460-464 */
            if ((double)topScore < o->score_percent_req * totalScore || topQual < o->moderate_quality)
                needToCheckRef = 1;
        }
        if (topScore < o->base_score_req || topQual <= o->low_quality) needToCheckRef = 1;

        if (needToCheckRef && refbase != 0) { /* group.cpp:470-501 */
            uint8_t refbase4bit = base2fourbits(refbase);
            signed char refBaseQual = 0; /* `char` in the reference */
            for (int r = 0; r < nreads; r++) {
                int readpos = i;
                if (!isLeft) readpos = i + lenDiff[r];
                if (readpos < 0 || readpos >= RD(w, reads[r])->l_qseq) continue;
                uint8_t qual = QUAL(w, reads[r])[readpos];
                uint8_t base = base_at(SEQ(w, reads[r]), readpos);
                if (base == refbase4bit) {
                    if ((int)qual > (int)refBaseQual) refBaseQual = (signed char)qual;
                    if (qual >= o->high_quality) topBase = refbase4bit;
                }
            }
            if (topQual < o->moderate_quality) topBase = refbase4bit;
            if (topBase == refbase4bit) topQual = (uint8_t)refBaseQual;
        }
        uint8_t outBase = base_at(outdata, i);
        if (outBase != topBase) { /* group.cpp:509-524 */
            if (i % 2 == 1) outdata[i / 2] = (uint8_t)((outdata[i / 2] & 0xF0) | topBase);
            else outdata[i / 2] = (uint8_t)((outdata[i / 2] & 0x0F) | (topBase << 4));
            diff++;
            if (refbase != 0) {
                uint8_t refbase4bit = base2fourbits(refbase);
                if (outBase == refbase4bit) mismatchInc++;
                else if (topBase == refbase4bit) mismatchInc--;
            }
        }
        outqual[i] = topQual;
    }
    if (mismatchInc > 5) { /* group.cpp:538-566: rollback; NM patch (568-572) is the caller's */
        memcpy(SEQ(w, out), seqBak, seqbytes);
        memcpy(QUAL(w, out), qualBak, qualbytes);
    }
    free(seqBak);
    free(qualBak);
    free(lenDiff);
    *mismatch_inc_out = mismatchInc;
    return diff;
}

static char fourbits2base(uint8_t v) { /* bamutil.cpp:149-165 */
    switch (v) {
        case 1: return 'A';
        case 2: return 'C';
        case 4: return 'G';
        case 8: return 'T';
        default: return 'N';
    }
}

/* getCigar string equality (bamutil.cpp:191-202): op chars + lengths */
static int same_cigar_string(const W *w, int a, int b) {
    static const char *OPS = "MIDNSHP=XB??????";
    if (RD(w, a)->n_cigar != RD(w, b)->n_cigar) return 0;
    const uint32_t *ca = CIG(w, a), *cb = CIG(w, b);
    for (int k = 0; k < RD(w, a)->n_cigar; k++)
        if (OPS[OP(ca[k])] != OPS[OP(cb[k])] || OPLEN(ca[k]) != OPLEN(cb[k])) return 0;
    return 1;
}

/* Group::consensusMergeBam, group.cpp:136-318. pairs[0..n) = the group's pairs in map order.
 * Returns the template's read slot or -1. */
static int consensus_merge_bam(W *w, const int *pairs, int n, int isLeft, int contig, int *diff, int *mismatch_inc) {
    const gcb_options *o = w->opt;
    int side = isLeft ? 0 : 1;
#define SLOT(k) (2 * pairs[k] + side)
#define HAVE(k) (w->have[SLOT(k)])
    if (n > o->skip_low_complexity_cluster_threshold) { /* group.cpp:142-175 */
        int distinct = 0, firstRead = -1;
        for (int k = 0; k < n; k++) {
            if (!HAVE(k)) continue;
            int seen = 0;
            for (int j = 0; j < k && !seen; j++)
                if (HAVE(j) && same_cigar_string(w, SLOT(j), SLOT(k))) seen = 1;
            if (!seen) distinct++;
            if (firstRead < 0) firstRead = SLOT(k);
        }
        if ((double)distinct > n * 0.1 && firstRead >= 0) {
            int sl = RD(w, firstRead)->l_qseq, diffNeighbor = 0;
            const uint8_t *s = SEQ(w, firstRead);
            for (int i = 0; i < sl - 1; i++)
                if (fourbits2base(base_at(s, i)) != fourbits2base(base_at(s, i + 1))) diffNeighbor++;
            if ((double)diffNeighbor < sl * 0.5) return -1;
        }
    }
    int leftReadMode = isLeft;
    if (!isLeft) { /* group.cpp:177-194 */
        int leftAligned = 1, lastPos = -1;
        for (int k = 0; k < n; k++) {
            if (HAVE(k)) {
                if (lastPos >= 0 && RD(w, SLOT(k))->pos != lastPos) { leftAligned = 0; break; }
                lastPos = RD(w, SLOT(k))->pos;
            }
        }
        if (leftAligned) leftReadMode = 1;
    }
    int *containedByList = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++) { /* group.cpp:196-233 */
        if (!HAVE(i)) continue;
        int containedBy = 1;
        for (int j = 0; j < n; j++) {
            if (i == j || !HAVE(j)) continue;
            if (!isLeft && get_right_ref_pos(w, SLOT(i)) != get_right_ref_pos(w, SLOT(j))) continue;
            if (is_part_of_slots(w, SLOT(i), SLOT(j), leftReadMode)) containedBy++;
        }
        containedByList[i] = containedBy;
        if (n > o->skip_low_complexity_cluster_threshold && containedBy >= n / 2) break;
    }
    int mostContainedById = -1, mostContainedByNum = -1;
    for (int i = 0; i < n; i++) { /* group.cpp:235-261 */
        if (containedByList[i] > mostContainedByNum) {
            mostContainedByNum = containedByList[i];
            mostContainedById = i;
        } else if (containedByList[i] == mostContainedByNum && mostContainedById >= 0) {
            int thisLen = HAVE(i) ? RD(w, SLOT(i))->l_qseq : 0;
            int curLen = HAVE(mostContainedById) ? RD(w, SLOT(mostContainedById))->l_qseq : 0;
            if (thisLen < curLen) {
                mostContainedByNum = containedByList[i];
                mostContainedById = i;
            }
        }
    }
    free(containedByList);
    if ((double)mostContainedByNum < n * 0.4 && n != 1) return -1; /* group.cpp:264 */

    int out = -1;
    signed char *outScore = get_score(w, pairs[mostContainedById], side); /* group.cpp:270-281 */
    if (HAVE(mostContainedById)) {
        out = SLOT(mostContainedById);
        w->have[out] = 0; /* stolen from its Pair */
    }
    if (out < 0) return -1;

    int *reads = (int *)malloc(sizeof(int) * n);
    signed char **scores = (signed char **)malloc(sizeof(signed char *) * n);
    int nreads = 0;
    reads[nreads] = out;
    scores[nreads++] = outScore;
    for (int j = 0; j < n; j++) { /* group.cpp:293-313 */
        if (j == mostContainedById) continue;
        signed char *score = get_score(w, pairs[j], side); /* evaluated even when the read is NULL */
        if (!HAVE(j) || !score) continue;
        if (is_part_of_slots(w, out, SLOT(j), leftReadMode)) {
            reads[nreads] = SLOT(j);
            scores[nreads++] = score;
        }
    }
    *diff = make_consensus(w, reads, scores, nreads, out, leftReadMode, contig, mismatch_inc);
    free(reads);
    free(scores);
    return out;
#undef SLOT
#undef HAVE
}

/* Cluster::duplexMergeBam, cluster.cpp:200-244, on two consensus records of the working copy */
static int duplex_merge_bam(W *w, int s1, int s2) {
    int len1 = RD(w, s1)->l_qseq, len2 = RD(w, s2)->l_qseq;
    int diff = abs(len1 - len2);
    int len = imin(len1, len2);
    uint8_t *seq1 = SEQ(w, s1), *seq2 = SEQ(w, s2), *qual1 = QUAL(w, s1), *qual2 = QUAL(w, s2);
    const uint8_t N4bits = 15;
    for (int i = 0; i < len; i++) {
        if (seq1[i / 2] == seq2[i / 2]) { /* the byte shortcut with its index-parity side effect */
            i++;
            continue;
        }
        char base1, base2;
        if (i % 2 == 1) {
            base1 = fourbits2base(seq1[i / 2] & 0xF);
            base2 = fourbits2base(seq2[i / 2] & 0xF);
        } else {
            base1 = fourbits2base((seq1[i / 2] >> 4) & 0xF);
            base2 = fourbits2base((seq2[i / 2] >> 4) & 0xF);
        }
        if (base1 != base2) {
            diff++;
            qual1[i] = 0;
            qual2[i] = 0;
            if (i % 2 == 1) {
                seq1[i / 2] = (uint8_t)((seq1[i / 2] & 0xF0) | N4bits);
                seq2[i / 2] = (uint8_t)((seq2[i / 2] & 0xF0) | N4bits);
            } else {
                seq1[i / 2] = (uint8_t)((seq1[i / 2] & 0x0F) | (N4bits << 4));
                seq2[i / 2] = (uint8_t)((seq2[i / 2] & 0x0F) | (N4bits << 4));
            }
        }
    }
    return diff;
}

/* field-wise compare of two UMI codes == std::string compare of the UMIs */
static int umi_cmp(const uint64_t *a, const uint64_t *b, int nw) {
    for (int k = 0; k < 16 * nw; k++) {
        int fa = (int)((a[k >> 4] >> (60 - 4 * (k & 15))) & 0xF), fb = (int)((b[k >> 4] >> (60 - 4 * (k & 15))) & 0xF);
        if (fa != fb) return fa < fb ? -1 : 1;
    }
    return 0;
}

/* Cluster::clusterByUMI, cluster.cpp:55-188, for cluster c */
static void cluster_by_umi(W *w, int c, gcb_result *res, int64_t *out_cursor, int *status) {
    const gcb_batch *b = w->b;
    const gcb_options *o = w->opt;
    int p0 = b->cluster_pair_off[c], p1 = b->cluster_pair_off[c + 1], n = p1 - p0;
    int nw = b->umi_words;
    int thr = b->cluster_flags[c] >> GCB_CLUSTER_UMI_THR_SHIFT;
    int crossContig = b->cluster_flags[c] & GCB_CLUSTER_CROSS_CONTIG;
    int contig = b->cluster_ref[c];
    char umibuf1[16 * GCB_MAX_UMI_WORDS + 1], umibuf2[16 * GCB_MAX_UMI_WORDS + 1];

    /* cluster.cpp:57-65: umiCount as a std::map<string,int>; here: count per pair of its UMI's multiplicity,
       kept per distinct UMI through a representative pair */
    int *count = (int *)calloc(n, sizeof(int)); /* count[i] valid on the first pair carrying each distinct UMI */
    int *rep = (int *)malloc(sizeof(int) * n);  /* representative (first pair) of pair i's UMI */
    int *remaining = (int *)malloc(sizeof(int) * n);
    int hasUMI = 0;
    for (int i = 0; i < n; i++) {
        const uint64_t *u = b->umi + (size_t)(p0 + i) * nw;
        if (u[0] >> 60) hasUMI = 1;
        rep[i] = i;
        for (int j = 0; j < i; j++)
            if (umi_cmp(u, b->umi + (size_t)(p0 + j) * nw, nw) == 0) { rep[i] = rep[j]; break; }
        count[rep[i]]++;
        remaining[i] = 1;
        res->pair_group[p0 + i] = -1;
    }
    int left_cnt = n, ngroups = 0;
    while (left_cnt > 0) { /* cluster.cpp:66-100 */
        int top = -1, topCount = 0;
        /* iterate distinct UMIs in std::map (lexicographic) order, strict > keeps the first maximum */
        for (int i = 0; i < n; i++) {
            if (rep[i] != i) continue;
            int better = 0;
            if (count[i] > topCount) better = 1;
            else if (count[i] == topCount && top >= 0 && count[i] > 0 &&
                     umi_cmp(b->umi + (size_t)(p0 + i) * nw, b->umi + (size_t)(p0 + top) * nw, nw) < 0)
                better = 1;
            if (better) { top = i; topCount = count[i]; }
        }
        const uint64_t *topUMI = top >= 0 ? b->umi + (size_t)(p0 + top) * nw : NULL;
        uint64_t empty[GCB_MAX_UMI_WORDS] = {0};
        if (!topUMI) topUMI = empty; /* cannot happen (every remaining pair keeps a positive count) */
        decode_umi(topUMI, nw, umibuf1);
        for (int i = 0; i < n; i++) {
            if (!remaining[i]) continue;
            decode_umi(b->umi + (size_t)(p0 + i) * nw, nw, umibuf2);
            if (gco_umi_diff(umibuf2, umibuf1) <= thr) {
                res->pair_group[p0 + i] = ngroups;
                remaining[i] = 0;
                left_cnt--;
                count[rep[i]] = 0;
            }
        }
        if (top >= 0) count[top] = 0;
        ngroups++;
    }
    res->cluster_n_groups[c] = ngroups;

    /* cluster.cpp:109-114: Group::consensusMerge per group (group.cpp:68-134) */
    int *gp = (int *)malloc(sizeof(int) * n);
    for (int g = 0; g < ngroups; g++) {
        gcb_group_result *gr = &res->groups[p0 + g];
        memset(gr, 0, sizeof *gr);
        gr->tmpl_read[0] = gr->tmpl_read[1] = -1;
        gr->qname_donor[0] = gr->qname_donor[1] = -1;
        gr->out_off[0] = gr->out_off[1] = -1;
        gr->duplex_partner = -1;
        gr->umi_pair = -1;
        int m = 0;
        for (int i = 0; i < n; i++)
            if (res->pair_group[p0 + i] == g) gp[m++] = p0 + i;
        if (m == 1 && RD(w, 2 * gp[0] + 1)->l_qseq < 0) { /* group.cpp:73-77 */
            gr->merge_reads = 1; /* Pair::Pair, pair.cpp:10 */
            gr->tmpl_read[0] = RD(w, 2 * gp[0])->l_qseq >= 0 ? 2 * gp[0] : -1;
            gr->umi_pair = gp[0];
            continue;
        }
        int nameToCopy = -1; /* group.cpp:79-99 */
        if (crossContig) {
            int curLen = 0;
            for (int k = 0; k < m; k++) {
                int sl = 2 * gp[k];
                if (RD(w, sl)->l_qseq < 0) continue;
                if (nameToCopy < 0) { nameToCopy = sl; curLen = RD(w, sl)->l_qname; continue; }
                /* equal padded length + strcmp < 0 cannot select a later map entry: map order IS strcmp order */
                if (RD(w, sl)->l_qname < curLen) { nameToCopy = sl; curLen = RD(w, sl)->l_qname; }
            }
        }
        int left = consensus_merge_bam(w, gp, m, 1, contig, &gr->diff[0], &gr->mismatch_inc[0]);
        int right = consensus_merge_bam(w, gp, m, 0, contig, &gr->diff[1], &gr->mismatch_inc[1]);
        gr->merge_reads = m; /* group.cpp:105 */
        gr->tmpl_read[0] = left;
        gr->tmpl_read[1] = right;
        int name_slot = -1;
        if (crossContig) { /* group.cpp:109-113 */
            if (left >= 0 && nameToCopy >= 0 && nameToCopy != left) gr->qname_donor[0] = nameToCopy;
            name_slot = left >= 0 ? (nameToCopy >= 0 ? nameToCopy : left) : right;
        } else if (left >= 0 && right >= 0) { /* group.cpp:114-123 */
            if (RD(w, left)->l_qname <= RD(w, right)->l_qname) { gr->qname_donor[1] = left; name_slot = left; }
            else { gr->qname_donor[0] = right; name_slot = right; }
        } else
            name_slot = left >= 0 ? left : right;
        gr->umi_pair = name_slot >= 0 ? name_slot / 2 : -1; /* Pair::setLeft/setRight, pair.cpp:188-216 */
    }
    free(gp);

    /* cluster.cpp:116-183 */
    int *alive = (int *)malloc(sizeof(int) * (ngroups > 0 ? ngroups : 1)); /* singleConsensusPairs as a stack */
    int nalive = ngroups;
    for (int g = 0; g < ngroups; g++) alive[g] = g;
    uint64_t zero[GCB_MAX_UMI_WORDS] = {0};
#define GUMI(g) (res->groups[p0 + (g)].umi_pair >= 0 ? b->umi + (size_t)res->groups[p0 + (g)].umi_pair * nw : zero)
    if (hasUMI && !o->disable_duplex) {
        while (nalive > 0) {
            int g1 = alive[--nalive];
            gcb_group_result *r1 = &res->groups[p0 + g1];
            decode_umi(GUMI(g1), nw, umibuf1);
            int found = 0;
            for (int i = 0; i < nalive; i++) {
                int g2 = alive[i];
                gcb_group_result *r2 = &res->groups[p0 + g2];
                decode_umi(GUMI(g2), nw, umibuf2);
                if (gco_is_duplex(umibuf1, umibuf2)) {
                    found = 1;
                    int diff = 0; /* Cluster::duplexMerge, cluster.cpp:190-198 */
                    if (r1->tmpl_read[0] >= 0 && r2->tmpl_read[0] >= 0) diff += duplex_merge_bam(w, r1->tmpl_read[0], r2->tmpl_read[0]);
                    if (r1->tmpl_read[1] >= 0 && r2->tmpl_read[1] >= 0) diff += duplex_merge_bam(w, r1->tmpl_read[1], r2->tmpl_read[1]);
                    r1->duplex_partner = g2;
                    r1->duplex_diff = diff;
                    r2->duplex_partner = g1;
                    r2->duplex_diff = diff;
                    r2->status = GCB_GROUP_DUPLEX_PARTNER;
                    if (diff <= o->duplex_mismatch_threshold) {
                        if (r1->merge_reads + r2->merge_reads >= o->cluster_size_req) {
                            r1->status = GCB_GROUP_DCS;
                            r1->reverse_merge_reads = r2->merge_reads; /* Pair::setDuplex */
                        } else
                            r1->status = GCB_GROUP_DUPLEX_SMALL;
                    } else
                        r1->status = GCB_GROUP_DUPLEX_DIFF;
                    for (int k = i; k + 1 < nalive; k++) alive[k] = alive[k + 1];
                    nalive--;
                    break;
                }
            }
            if (!found) {
                if (!o->duplex_only && r1->merge_reads >= o->cluster_size_req) r1->status = GCB_GROUP_SSCS;
                else r1->status = GCB_GROUP_DROPPED;
            }
        }
    } else {
        for (int g = 0; g < ngroups; g++) {
            gcb_group_result *r = &res->groups[p0 + g];
            if (!o->duplex_only && r->merge_reads >= o->cluster_size_req) r->status = GCB_GROUP_SSCS;
            else r->status = GCB_GROUP_DROPPED;
        }
    }
#undef GUMI
    free(alive);

    /* lay the consensus records out: cluster order, group order, side order */
    for (int g = 0; g < ngroups; g++) {
        gcb_group_result *gr = &res->groups[p0 + g];
        for (int s = 0; s < 2; s++) {
            int t = gr->tmpl_read[s];
            if (t < 0) continue;
            int l = RD(w, t)->l_qseq;
            int64_t sz = GCB_ALIGN4(l) + GCB_ALIGN4((l + 1) / 2);
            if (*out_cursor + sz > res->out_capacity) { *status = GCB_ERR_CAPACITY; continue; }
            gr->out_off[s] = *out_cursor;
            memset(res->out_payload + *out_cursor, 0, (size_t)sz);
            memcpy(res->out_payload + *out_cursor, QUAL(w, t), l);
            memcpy(res->out_payload + *out_cursor + GCB_ALIGN4(l), SEQ(w, t), (l + 1) / 2);
            *out_cursor += sz;
        }
    }
    free(count);
    free(rep);
    free(remaining);
}

int gco_consensus_batch(const gcb_options *opt, const gco_genome *genome, const gcb_batch *batch, gcb_result *result) {
    if (!opt || !batch || !result) return GCB_ERR_ARG;
    if (batch->umi_words < 1 || batch->umi_words > GCB_MAX_UMI_WORDS) return GCB_ERR_ARG;
    W w;
    w.opt = opt;
    w.genome = genome;
    w.b = batch;
    w.work = (uint8_t *)malloc(batch->payload_bytes > 0 ? (size_t)batch->payload_bytes : 1);
    memcpy(w.work, batch->payload, (size_t)batch->payload_bytes);
    size_t ns = 2 * (size_t)batch->n_pairs;
    w.score = (signed char **)calloc(ns ? ns : 1, sizeof(signed char *));
    w.have = (int32_t *)malloc(sizeof(int32_t) * (ns ? ns : 1));
    for (size_t s = 0; s < ns; s++) w.have[s] = batch->reads[s].l_qseq >= 0;
    memset(result->groups, 0, sizeof(gcb_group_result) * (size_t)batch->n_pairs);
    int status = GCB_OK;
    int64_t cursor = 0;
    for (int c = 0; c < batch->n_clusters; c++) cluster_by_umi(&w, c, result, &cursor, &status);
    *result->out_bytes = cursor;
    for (size_t s = 0; s < ns; s++) free(w.score[s]);
    free(w.score);
    free(w.have);
    free(w.work);
    return status;
}
