// vote_column.cuh — the per-column decision of Group::makeConsensus (group.cpp:395-525) as a pure
// function of the column's histogram, so that every kernel variant shares one statement of the
// reference's tie-breaking and reference-arbitration rules.
#pragma once

#include "device_common.cuh"

namespace gcb {

// One observed base code of a column: what group.cpp:376-393 accumulates in counts[]/baseScores[]/quals[]/topQuals[]
struct VoteBin {
    int base;   // 4-bit code 0..15
    int cnt;    // counts[base]
    int score;  // baseScores[base]
    int qual;   // quals[base]      (a SUM)
    int maxq;   // topQuals[base]
};

// (score, qual, base) lexicographic: the scans at group.cpp:395-402 and 407-416 walk b = 0..15 and
// replace on `>` score or equal score and `>=` qual, i.e. they return the lexicographic maximum
// with ties going to the LARGER code; bins nobody voted for take part with (0, 0).
GCB_HD bool bin_beats(const VoteBin &a, const VoteBin &b) {
    if (a.score != b.score) return a.score > b.score;
    if (a.qual != b.qual) return a.qual > b.qual;
    return a.base > b.base;
}

struct ColumnTop {
    VoteBin top, sec;
    bool fast;           // group.cpp:421-427: only the quality is written, the base is left alone
    bool need_ref;       // needToCheckRef after group.cpp:421-467
};

// group.cpp:419-467: what the top and second bins imply.  `r.top` / `r.sec` are filled in; sets r.fast / r.need_ref.
GCB_HD void column_rules(const gcb_options &o, ColumnTop &r, int total) {
    const int topScore = r.top.score, topNum = r.top.cnt, topQual = r.top.maxq, secNum = r.sec.cnt;
    r.fast = false;
    r.need_ref = false;
    if (secNum == 0) {
        if (topScore >= o.base_score_req && topQual >= o.moderate_quality) {
            r.fast = true;
            return;
        }
        r.need_ref = true;
    }
    if (secNum == 1) {  // group.cpp:442-457; quals[secBase] is a sum of one quality here
        if (r.sec.qual <= o.low_quality) {
            if (topNum < 2 && topQual < o.high_quality) r.need_ref = true;
        } else {
            if (topNum < 3 || topQual < o.high_quality) r.need_ref = true;
        }
    }
    if (secNum > 1) {  // group.cpp:460-464
        if ((double)topScore < o.score_percent_req * (double)total || topQual < o.moderate_quality) r.need_ref = true;
    }
    if (topScore < o.base_score_req || topQual <= o.low_quality) r.need_ref = true;
}

// group.cpp:395-467.  obs[0..nobs) are the distinct observed codes (cnt >= 1), total = totalScore.
GCB_HD ColumnTop column_top(const gcb_options &o, const VoteBin *obs, int nobs, int total) {
    // the two largest codes nobody voted for
    unsigned freemask = 0xFFFFu;
    for (int k = 0; k < nobs; k++) freemask &= ~(1u << obs[k].base);
    VoteBin e1 = {-1, 0, 0, 0, 0}, e2 = {-1, 0, 0, 0, 0};
    for (int b = 15; b >= 0; b--) {
        if (!((freemask >> b) & 1u)) continue;
        if (e1.base < 0) e1.base = b;
        else { e2.base = b; break; }
    }
    ColumnTop r;
    VoteBin none = {-1, 0, -0x7FFFFFFF, -1, 0};
    r.top = none;
    int topk = -1;  // index in obs, or -2/-3 for e1/e2
    for (int k = 0; k < nobs; k++)
        if (bin_beats(obs[k], r.top)) { r.top = obs[k]; topk = k; }
    if (e1.base >= 0 && bin_beats(e1, r.top)) { r.top = e1; topk = -2; }
    // e2 can never beat e1 (same score and qual, smaller code)
    r.sec = none;
    for (int k = 0; k < nobs; k++)
        if (k != topk && bin_beats(obs[k], r.sec)) r.sec = obs[k];
    if (topk != -2 && e1.base >= 0 && bin_beats(e1, r.sec)) r.sec = e1;
    if (topk == -2 && e2.base >= 0 && bin_beats(e2, r.sec)) r.sec = e2;
    // sixteen codes all observed and only one of them... cannot leave sec unset: nobs + free codes = 16 >= 2

    column_rules(o, r, total);
    return r;
}

struct ColumnOut {
    int base;        // topBase to compare with the template's base (group.cpp:509)
    int qual;        // topQual written at group.cpp:525
};

// group.cpp:470-501.  ref4 = BAM code of the reference base (1,2,4,8) or 0 when there is none.
// ref_max_qual / ref_any_high summarise the reads whose base equals ref4: the `char refBaseQual`
// running maximum (as the reference computes it) and whether one of them has qual >= highQuality.
GCB_HD ColumnOut column_arbitrate(const gcb_options &o, const ColumnTop &t, int ref4, int ref_base_qual_char, bool ref_any_high) {
    ColumnOut out;
    out.base = t.top.base;
    out.qual = t.top.maxq;
    if (t.need_ref && ref4 != 0) {
        if (ref_any_high) out.base = ref4;
        if (out.qual < o.moderate_quality) out.base = ref4;
        if (out.base == ref4) out.qual = (int)(uint8_t)ref_base_qual_char;
    }
    return out;
}

// FastaReader::getBase + bits2base (fastareader.cpp:115-128) followed by BamUtil::base2fourbits
// (bamutil.cpp:167-183): genome nibble (A=1 T=2 C=3 G=4) -> BAM code, 0 = no usable reference base
GCB_HD int genome_nibble_to_bam(int bits) {
    return bits == 1 ? 1 : bits == 2 ? 8 : bits == 3 ? 2 : bits == 4 ? 4 : 0;
}

}  // namespace gcb
