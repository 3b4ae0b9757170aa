// k_slow_columns.cuh — the slow columns of the vote: group.cpp:376-525 for every column the ring kernel (k_vote_ring.cuh)
// could not finish in the word, one THREAD per column.
//
//   finish_column()   the decision once the column's histogram is known: top-2 selection (group.cpp:395-417), the rules
//                     and the reference arbitration (group.cpp:419-525); patches the consensus record and adds to the family
//                     side's diff / mismatchInc (atomics on the result row, which select_template_kernel wrote with zeros).
//   decide_column()   histogram straight from the family side's VoteRead entries and its cluster's slab (score per read:
//                     pair.cpp:121-170; three-bin register histogram: group.cpp:376-393).  The ring kernel's voter warps
//                     call it for tiles of deep families, from the staged slab in shared memory.
//   slow_columns_kernel   everything else.  The ring kernel keeps a tile in shared memory only as long as its warps vote;
//                     what a slow column needs of the tile — per read its quality, its base nibble, its mate's quality and
//                     base nibble and where pair.cpp:121-170 puts the column: 4 bytes per read behind a 32-byte
//                     self-contained header — is written by the lane that found the column into a global queue (one
//                     64-bit atomic per ~10 bundles reserves a warp's records and words in one counter), and this kernel
//                     takes one record per thread at full occupancy: the deciding is a chain of dependent small loads that
//                     wants many resident warps, which the one-CTA-per-SM ring cannot give it.  (Measured alternatives,
//                     profiles/r03_notes.md: deciding inside the ring — from the staged slab, or from the L2 by dedicated
//                     warps —, extracting per tile instead of per bundle, re-reading the payload from a second kernel.)
#pragma once

#include "vote_tile.cuh"

namespace gcb {

constexpr uint32_t VQ_INVALID = 0xFFFFFFFFu;   // index entry of a reservation that was not used
constexpr int VQ_SLOW_THREADS = 128;
constexpr int VQ_SLOW_CTAS = 148 * 8;          // slow_columns_kernel strides over the records
constexpr uint32_t VQ_POOL_RECS = 64, VQ_POOL_WORDS = 64 * 20;  // queue space a voter warp reserves at a time

struct SlowQueue {
    unsigned long long *count;   // [1] records << 32 | words reserved so far (may run past the capacity)
    uint32_t *words;             // [cap_words] records (see SR_HDR_WORDS)
    uint32_t *index;             // [cap_recs] word offset of every record, VQ_INVALID = none
    uint32_t cap_words, cap_recs;
};

// record: SR_HDR_WORDS header words, then n entries (one per read of the family side), padded to a multiple of 4 words.
// The header is self-contained (slow_columns_kernel needs no table lookup):
//   [0] 2 * slot + side   [1] col | n << 16   [2] tmpl_k | flags << 16   [3] l_out
//   [4..5] absolute offset of the consensus record in out_payload   [6..7] ref_nib0
constexpr int SR_HDR_WORDS = 8;
constexpr uint32_t SR_UNVOTED = 0x100u;    // (next to the FS_* flags) column beyond the voted length: the record keeps the template's (rewritten) quality
// entry: quality | mate quality << 8 | base << 16 | mate base << 20 | state << 24 | SE_VOTES
constexpr uint32_t SE_VOTES = 1u << 26;
constexpr uint32_t SE_NO_INFO = 0u, SE_PLAIN = 1u, SE_MATE = 2u, SE_NO_MATE_BASE = 3u;

GCB_DEV uint32_t slow_rec_words(int m) { return (uint32_t)SR_HDR_WORDS + (((uint32_t)m + 3u) & ~3u); }
GCB_DEV void slow_write_header(uint32_t *rec, const FsTile &ft, int col, int64_t out_abs) {
    const uint32_t side = (ft.flags & FS_SIDE1) ? 1u : 0u;
    const uint32_t fl = (uint32_t)ft.flags | (col >= (int)ft.len ? SR_UNVOTED : 0u);
    uint4 a, c;
    a.x = 2u * (uint32_t)ft.slot + side;
    a.y = (uint32_t)col | ((uint32_t)ft.m << 16);
    a.z = (uint32_t)ft.tmpl_k | (fl << 16);
    a.w = (uint32_t)ft.l_out;
    c.x = (uint32_t)(uint64_t)out_abs; c.y = (uint32_t)((uint64_t)out_abs >> 32);
    c.z = (uint32_t)(uint64_t)ft.ref_nib0; c.w = (uint32_t)((uint64_t)ft.ref_nib0 >> 32);
    ((uint4 *)rec)[0] = a;
    ((uint4 *)rec)[1] = c;
}

// what pair.cpp:88-172 needs of read `v` at template column `col`, as a queue entry (0 = the read has no base there)
GCB_DEV uint32_t slow_entry(const uint8_t *cb, const VoteRead &v, int col) {
    const int rp = col + v.shift;
    if (v.own_off4 == VR_NO_VOTE || rp < 0 || rp >= v.own_l) return 0u;
    const uint8_t *q = cb + 4 * (int)v.own_off4;
    const uint32_t ql = q[rp];
    const uint32_t base = (uint32_t)base_at(q + GCB_ALIGN4(v.own_l), rp);
    const bool info = v.ov_len != VR_NO_OVERLAP_INFO;
    const int k = rp - v.ov_own, mp = v.ov_mate + k;
    const bool inwin = info && k >= 0 && k < v.ov_len;
    const bool mvalid = inwin && mp >= 0 && mp < v.mate_l;
    uint32_t mql = 0u, mbase = 0u;
    if (mvalid) {
        const uint8_t *mq = cb + 4 * (int)v.mate_off4;
        mql = mq[mp];
        mbase = (uint32_t)base_at(mq + GCB_ALIGN4(v.mate_l), mp);
    }
    const uint32_t st = !info ? SE_NO_INFO : !inwin ? SE_PLAIN : mvalid ? SE_MATE : SE_NO_MATE_BASE;
    return ql | (mql << 8) | (base << 16) | (mbase << 20) | (st << 24) | SE_VOTES;
}

// base, rewritten quality and score of a queue entry: the same function of the same bytes as fetch_vote
GCB_DEV bool slow_decode(const ScoreTab &t, uint32_t ent, int side, int &base, int &qual, int &score) {
    if (!(ent & SE_VOTES)) return false;
    const int ql = (int)(ent & 0xFFu), mql = (int)((ent >> 8) & 0xFFu);
    base = (int)((ent >> 16) & 0xFu);
    const int mbase = (int)((ent >> 20) & 0xFu);
    const uint32_t st = (ent >> 24) & 3u;
    // pair.cpp:121-170 with one qual2score: of the read's quality outside the overlap, of the mean quality when the mates agree
    // (+4), of the difference when they do not (-3 for the better read — the left one on a tie —, nothing for the other, whose
    // quality is rewritten)
    const bool mate = st == SE_MATE, match = base == mbase, ge = ql >= mql;
    const bool mine = side == 0 ? ge : (mql < ql);
    const int sc = t.q2s(!mate ? ql : match ? (ql + mql) >> 1 : ge ? ql - mql : mql - ql);
    score = mate ? (match ? sc8(sc + 4) : mine ? sc8(sc - 3) : 0) : st == SE_PLAIN ? sc : t.sm;
    qual = (mate && !match) ? max(0, ql - mql) : ql;
    return true;
}

// what a column's decision needs to know of its family side
struct SlowSide {
    int m, l_out, len, tmpl_k, side, flags, slot;
    int64_t ref_nib0;
};

// group.cpp:395-525 for one column whose three-bin histogram is `bins`.  `each(f)` calls f(base, quality, score, is_template)
// for every voting read of the column in map order (only the rare paths walk the reads again: a fourth distinct code, and
// a reference-agreeing quality of 128 or more, where the scan order matters).  obase: the record's base before the vote.
template <typename Each>
GCB_DEV void finish_column(const BatchView &b, const ResultView &r, const GenomeView &gv, const gcb_options &o, const RollbackList &rb,
                           const SlowSide &fs, uint8_t *out, int col, const Bins3 &bins, int obase, Each &&each) {
    const int side = fs.side, qbytes = GCB_ALIGN4(fs.l_out);
    ColumnTop ct;
    int total = bins.total;
    uint32_t acgt = 0;  // the best quality of codes 1, 2, 4, 8 in bytes 0..3 (0 when nobody showed the code)
    if (bins.overflow) {  // a fourth distinct code: the sixteen-bin histogram in local memory (group.cpp:376-417 as written)
        int32_t h[64];
        for (int q = 0; q < 64; q++) h[q] = 0;
        each([&](int base, int qual, int score, bool) {
            h[4 * base]++;
            h[4 * base + 1] += score;
            h[4 * base + 2] += qual;
            h[4 * base + 3] = max(h[4 * base + 3], qual);
        });
        VoteBin obs[16];
        int nobs = 0;
        total = 0;
        for (int q = 0; q < 16; q++)
            if (h[4 * q] > 0) {
                obs[nobs].base = q; obs[nobs].cnt = h[4 * q]; obs[nobs].score = h[4 * q + 1]; obs[nobs].qual = h[4 * q + 2]; obs[nobs].maxq = h[4 * q + 3];
                total += obs[nobs].score;
                nobs++;
            }
        ct = column_top(o, obs, nobs, total);
        acgt = (uint32_t)(h[4 * 1] > 0 ? h[4 * 1 + 3] : 0) | ((uint32_t)(h[4 * 2] > 0 ? h[4 * 2 + 3] : 0) << 8) |
               ((uint32_t)(h[4 * 4] > 0 ? h[4 * 4 + 3] : 0) << 16) | ((uint32_t)(h[4 * 8] > 0 ? h[4 * 8 + 3] : 0) << 24);
    } else {
        // top and second: every bin competes with its (score, quality sum, code) key; the codes nobody showed compete
        // with (0, 0, code), of which only the two largest can place
        unsigned freemask = 0xFFFFu;
        unsigned long long key[3];
#pragma unroll
        for (int kk = 0; kk < 3; kk++) {
            const VoteBin vb = bins.bin(kk);
            const int bb = vb.base;
            const bool have = bb >= 0;
            key[kk] = have ? bin_key(vb.score, vb.qual, bb) : 0ull;
            if (have) freemask &= ~(1u << bb);
            if (have && (bb == 1 || bb == 2 || bb == 4 || bb == 8)) acgt |= (uint32_t)vb.maxq << (bb == 1 ? 0 : bb == 2 ? 8 : bb == 4 ? 16 : 24);
        }
        const int e1 = 31 - __clz((int)freemask);
        freemask &= ~(1u << e1);
        const int e2 = 31 - __clz((int)freemask);
        const unsigned long long ke1 = bin_key(0, 0, e1), ke2 = bin_key(0, 0, e2);
        unsigned long long top = max_u64(key[0], key[1]), sec = min_u64(key[0], key[1]);
        sec = max_u64(sec, min_u64(top, key[2])); top = max_u64(top, key[2]);
        sec = max_u64(sec, min_u64(top, ke1)); top = max_u64(top, ke1);
        sec = max_u64(sec, min_u64(top, ke2)); top = max_u64(top, ke2);
        const int tb = (int)(top & 0xF), sb = (int)(sec & 0xF);
        const VoteBin none = {0, 0, 0, 0, 0};
        ct.top = bins.b0 == tb ? bins.bin(0) : bins.b1 == tb ? bins.bin(1) : bins.b2 == tb ? bins.bin(2) : none;
        ct.sec = bins.b0 == sb ? bins.bin(0) : bins.b1 == sb ? bins.bin(1) : bins.b2 == sb ? bins.bin(2) : none;
        ct.top.base = tb;
        ct.sec.base = sb;
        column_rules(o, ct, total);
    }
    int new_qual;
    if (ct.fast) {
        new_qual = ct.top.maxq;  // group.cpp:422-426: the base is NOT written
    } else {
        int ref4 = 0;
        if (fs.flags & FS_REF_OK) {  // group.cpp:430-439
            int refpos = col;
            if (!(fs.flags & FS_SIMPLE_CIGAR)) {
                const gcb_read_desc od = b.reads[r.groups[fs.slot].tmpl_read[side]];
                refpos = get_ref_offset(b.cigar + od.cigar_off, od.n_cigar, col);
            }
            const int64_t nib = fs.ref_nib0 + refpos;
            if (refpos >= 0 && nib >= 0 && (nib >> 1) < gv.packed_bytes) {  // the bound only guards malformed CIGARs
                const uint8_t two = gv.packed4[nib >> 1];
                ref4 = genome_nibble_to_bam((nib & 1) ? (two >> 4) : (two & 0xF));
            }
        }
        int rbq = 0;
        bool any_high = false;
        if (ct.need_ref && ref4 != 0) {
            const int rmax = (int)((acgt >> (ref4 == 1 ? 0 : ref4 == 2 ? 8 : ref4 == 4 ? 16 : 24)) & 0xFFu);
            if (rmax >= 128) {  // `char refBaseQual` wraps: the scan order matters (group.cpp:474-490): template first
                each([&](int base, int qual, int, bool is_tmpl) {
                    if (!is_tmpl || base != ref4) return;
                    if (qual > rbq) rbq = sc8(qual);
                    if (qual >= o.high_quality) any_high = true;
                });
                each([&](int base, int qual, int, bool is_tmpl) {
                    if (is_tmpl || base != ref4) return;
                    if (qual > rbq) rbq = sc8(qual);
                    if (qual >= o.high_quality) any_high = true;
                });
            } else {
                rbq = rmax;
                any_high = rmax >= o.high_quality;
            }
        }
        const ColumnOut co = column_arbitrate(o, ct, ref4, rbq, any_high);
        if (obase != co.base) {  // group.cpp:509-524
            int d_mm = 0;
            if (ref4 != 0) {
                if (obase == ref4) d_mm = 1;
                else if (co.base == ref4) d_mm = -1;
            }
            gcb_group_result *gr = r.groups + fs.slot;
            atomicAdd(&gr->diff[side], 1);
            if (d_mm != 0) {
                const int before = atomicAdd(&gr->mismatch_inc[side], d_mm);
                if (d_mm > 0 && before == 5) {  // more than five new mismatches so far: vote_rollback_kernel looks at the final count
                    const int kk = atomicAdd(rb.count, 1);
                    if (kk < rb.cap) rb.list[kk] = 2 * fs.slot + side;
                }
            }
            const int byte = col >> 1;
            const unsigned delta = ((unsigned)(obase ^ co.base) & 0xFu) << ((col & 1) ? 0 : 4);
            atomicXor((unsigned *)(out + qbytes + (byte & ~3)), delta << (8 * (byte & 3)));
        }
        new_qual = co.qual;
    }
    out[col] = (uint8_t)new_qual;
}

#ifndef GCB_DC_UNROLL
#define GCB_DC_UNROLL 2  // (8: 2.20 ms, 4 and 2: 1.91 ms on the cfg4 shape: the smaller code wins)
#endif
// One column of a family side from its VoteRead entries `ents` and its cluster's slab `cb` (shared or global memory).
GCB_DEV void decide_column(const BatchView &b, const ResultView &r, const GenomeView &gv, const gcb_options &o, const RollbackList &rb,
                           const SlowSide &fs, const uint8_t *cb, const VoteRead *ents, uint8_t *out, int col) {
    const ScoreTab tab(o);
    const VoteRead tv = ents[fs.tmpl_k];
    const int side = fs.side;
    const int qbytes = GCB_ALIGN4(fs.l_out);
    GCB_COUNT(3, 1);
    if (col >= fs.len) {  // beyond the voted columns the record keeps what it held (rewritten qualities)
        int obase = 0, oqual = 0, sc;
        fetch_ent(cb, tv, col, side, o, obase, oqual, sc);
        out[col] = (uint8_t)oqual;
        return;
    }
    Bins3 bins;
    bins.init();
    const int obase = base_at(cb + 4 * (int)tv.own_off4 + qbytes, col);  // the record's base before the vote: the template's own
    bins.seed(obase);
    const int m = fs.m;
    bool general = true;
    if (fs.flags & FS_UNIFORM) {
        const bool info = tv.ov_len != VR_NO_OVERLAP_INFO;
        const int k = col - (int)tv.ov_own, mp = (int)tv.ov_mate + k;
        const bool inwin = info && k >= 0 && k < (int)tv.ov_len;
        const bool mvalid = inwin && mp >= 0 && mp < (int)tv.mate_l;
        const bool plain = info && !inwin;  // pair.cpp:121-131: outside the overlap the score follows the quality
        const int soff = qbytes + (col >> 1), nsh = (col & 1) ? 0 : 4;
        const int mpi = mvalid ? mp : 0;
        const int msoff = GCB_ALIGN4(tv.mate_l) + (mpi >> 1), mnsh = (mpi & 1) ? 0 : 4;
        // GCB_DC_UNROLL reads at a time, in two waves of independent loads (where their records lie, then their bytes).  ONE walk:
        // a second instantiation (two-bin histogram first, three-bin on overflow) doubles the code of a function every voter
        // warp of a deep tile enters, and measured 4.6 ms against 2.4 on the cfg4 shape; a two-bin walk that falls back to the
        // general loop 3.2 ms (with 50 reads at 1 % errors some lane of nearly every warp sees a third code): profiles/r03_notes.md.
        for (int e0 = 0; e0 < m; e0 += GCB_DC_UNROLL) {
            uint32_t w[GCB_DC_UNROLL], x[GCB_DC_UNROLL];
#pragma unroll
            for (int u = 0; u < GCB_DC_UNROLL; u++) w[u] = e0 + u < m ? *(const uint32_t *)(ents + e0 + u) : (uint32_t)VR_NO_VOTE;  // own_off4 | mate_off4 << 16
#pragma unroll
            for (int u = 0; u < GCB_DC_UNROLL; u++) {
                x[u] = 0u;
                if ((w[u] & 0xFFFFu) != VR_NO_VOTE) {
                    const uint8_t *rec = cb + 4 * (int)(w[u] & 0xFFFFu);
                    x[u] = (uint32_t)rec[col] | ((uint32_t)rec[soff] << 8);
                    if (mvalid) {
                        const uint8_t *mrec = cb + 4 * (int)(w[u] >> 16);
                        x[u] |= ((uint32_t)mrec[mpi] << 16) | ((uint32_t)mrec[msoff] << 24);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < GCB_DC_UNROLL; u++) {
                if ((w[u] & 0xFFFFu) == VR_NO_VOTE) continue;
                int ql = (int)(x[u] & 0xFFu);
                const int base = (int)((x[u] >> (8 + nsh)) & 0xFu);
                int score;
                if (mvalid) {
                    // pair.cpp:147-169 with one qual2score: of the mean quality when the mates agree (+4), of the difference
                    // when they do not (-3 for the better read, nothing for the other, whose quality is rewritten)
                    const int mql = (int)((x[u] >> 16) & 0xFFu);
                    const int mbase = (int)((x[u] >> (24 + mnsh)) & 0xFu);
                    const bool match = base == mbase, ge = ql >= mql;
                    const bool mine = side == 0 ? ge : (mql < ql);  // left read wins ties
                    const int sc = tab.q2s(match ? (ql + mql) >> 1 : ge ? ql - mql : mql - ql);
                    score = match ? sc8(sc + 4) : mine ? sc8(sc - 3) : 0;
                    ql = match ? ql : max(0, ql - mql);
                } else {
                    score = plain ? tab.q2s(ql) : tab.sm;
                }
                bins.add(base, ql, score);
            }
        }
        general = false;
    }
    if (general) {
        for (int e = 0; e < m; e++) {
            int base, qual, score;
            if (fetch_vote(cb, ents[e], col, side, tab, base, qual, score)) bins.add(base, qual, score);
        }
    }
    finish_column(b, r, gv, o, rb, fs, out, col, bins, obase, [&](auto &&f) {
        for (int e = 0; e < m; e++) {
            int base, qual, score;
            if (fetch_ent(cb, ents[e], col, side, o, base, qual, score)) f(base, qual, score, e == fs.tmpl_k);
        }
    });
}

// One queued column (group.cpp:376-525 from its record)
GCB_DEV void slow_record(const BatchView &b, const ResultView &r, const GenomeView &gv, const gcb_options &o, const ScoreTab &tab,
                         const RollbackList &rb, const uint32_t *rec) {
    const uint4 ha = ((const uint4 *)rec)[0], hc = ((const uint4 *)rec)[1];
    const uint32_t fsid = ha.x, w1 = ha.y, w2 = ha.z;
    const int col = (int)(w1 & 0xFFFFu), n = (int)(w1 >> 16);
    const uint32_t flags = w2 >> 16;
    const uint32_t *ents = rec + SR_HDR_WORDS;
    SlowSide fs;
    fs.m = n; fs.l_out = (int)ha.w; fs.len = (int)ha.w; fs.tmpl_k = (int)(w2 & 0xFFFFu); fs.side = (int)(fsid & 1u); fs.flags = (int)(flags & 0xFFu);
    fs.slot = (int)(fsid >> 1);
    fs.ref_nib0 = (int64_t)(((uint64_t)hc.w << 32) | hc.z);
    const int side = fs.side;
    uint8_t *out = r.out_payload + (int64_t)(((uint64_t)hc.y << 32) | hc.x);
    GCB_COUNT(3, 1);
    if (flags & SR_UNVOTED) {  // beyond the voted columns the record keeps what it held (rewritten qualities)
        int obase = 0, oqual = 0, sc;
        slow_decode(tab, ents[fs.tmpl_k], side, obase, oqual, sc);
        out[col] = (uint8_t)oqual;
        return;
    }
    // most slow columns show two codes: a two-bin histogram first, the three-bin one when a third code turns up
    Bins3 bins;
    bins.init();
    auto walk = [&](auto &acc) {
        for (int e = 0; e < n; e += 4) {  // (records are padded to whole 16-byte groups of entries)
            const uint4 v = *(const uint4 *)(ents + e);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int base, qual, score;
                if (e + k < n && slow_decode(tab, w[k], side, base, qual, score)) acc.add(base, qual, score);
            }
        }
    };
    const int obase = (int)((ents[fs.tmpl_k] >> 16) & 0xFu);  // the record's base before the vote: the template's own
    Bins2 two;
    two.init();
    two.seed(obase);
    walk(two);
    if (!two.overflow) {
        bins.b0 = two.bA; bins.c0 = two.cA; bins.s0 = two.sA; bins.q0 = two.qA; bins.x0 = two.xA;
        bins.b1 = two.bB; bins.c1 = two.cB; bins.s1 = two.sB; bins.q1 = two.qB; bins.x1 = two.xB;
        bins.total = two.total;
    } else {
        bins.seed(obase);
        walk(bins);
    }
    // the record's base before the vote: the template's own (pair.cpp rewrites qualities, never bases)
    finish_column(b, r, gv, o, rb, fs, out, col, bins, obase, [&](auto &&f) {
        for (int e = 0; e < n; e++) {
            int base, qual, score;
            if (slow_decode(tab, ents[e], side, base, qual, score)) f(base, qual, score, e == fs.tmpl_k);
        }
    });
}

__global__ void __launch_bounds__(VQ_SLOW_THREADS) slow_columns_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o,
                                                                       SlowQueue sq, RollbackList rb) {
    GCB_GRID_DEP();
    if (batch_is_malformed(ws.error_flag)) return;
    const uint32_t reserved = (uint32_t)(*sq.count >> 32), total = reserved < sq.cap_recs ? reserved : sq.cap_recs;
    const ScoreTab tab(o);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t off = sq.index[i];
        if (off == VQ_INVALID) continue;
        slow_record(b, r, gv, o, tab, rb, sq.words + off);
    }
}

}  // namespace gcb
