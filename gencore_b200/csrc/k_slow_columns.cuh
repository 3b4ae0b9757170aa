// k_slow_columns.cuh — the slow columns of the vote: group.cpp:376-525 for every column the ring kernel (k_vote_ring.cuh)
// could not finish in the word, one THREAD per column.
//
//   finish_column()   the decision once the column's histogram is known: top-2 selection (group.cpp:395-417), the rules
//                     and the reference arbitration (group.cpp:419-525); patches the consensus record and adds to the family
//                     side's diff / mismatchInc (atomics on the result row, which select_template_kernel wrote with zeros).
//   decide_column()   histogram straight from the family side's VoteRead entries and its cluster's slab (score per read:
//                     pair.cpp:121-170; three-bin register histogram: group.cpp:376-393).  The ring kernel's voter warps
//                     call it, 32 columns of a closed tile at a time, from the slab staged in shared memory.
#pragma once

#include "vote_tile.cuh"

namespace gcb {

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
    const int m = fs.m;
    if (fs.flags & FS_UNIFORM) {
        const bool info = tv.ov_len != VR_NO_OVERLAP_INFO;
        const int k = col - (int)tv.ov_own, mp = (int)tv.ov_mate + k;
        const bool inwin = info && k >= 0 && k < (int)tv.ov_len;
        const bool mvalid = inwin && mp >= 0 && mp < (int)tv.mate_l;
        const bool plain = info && !inwin;  // pair.cpp:121-131: outside the overlap the score follows the quality
        const int soff = qbytes + (col >> 1), nsh = (col & 1) ? 0 : 4;
        const int mpi = mvalid ? mp : 0;
        const int msoff = GCB_ALIGN4(tv.mate_l) + (mpi >> 1), mnsh = (mpi & 1) ? 0 : 4;
        // eight reads at a time, in two waves of independent loads (where their records lie, then their bytes)
        auto walk = [&](auto &acc) {
            for (int e0 = 0; e0 < m; e0 += 8) {
                uint32_t w[8], x[8];
#pragma unroll
                for (int u = 0; u < 8; u++) w[u] = e0 + u < m ? *(const uint32_t *)(ents + e0 + u) : (uint32_t)VR_NO_VOTE;  // own_off4 | mate_off4 << 16
#pragma unroll
                for (int u = 0; u < 8; u++) {
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
                for (int u = 0; u < 8; u++) {
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
                    acc.add(base, ql, score);
                }
            }
        };
        // most slow columns show two codes: a two-bin histogram first, the three-bin one when a third code turns up
        Bins2 two;
        two.init();
        walk(two);
        if (!two.overflow) {
            bins.b0 = two.bA; bins.c0 = two.cA; bins.s0 = two.sA; bins.q0 = two.qA; bins.x0 = two.xA;
            bins.b1 = two.bB; bins.c1 = two.cB; bins.s1 = two.sB; bins.q1 = two.qB; bins.x1 = two.xB;
            bins.total = two.total;
        } else {
            walk(bins);
        }
    } else {
        for (int e = 0; e < m; e++) {
            int base, qual, score;
            if (fetch_vote(cb, ents[e], col, side, tab, base, qual, score)) bins.add(base, qual, score);
        }
    }
    finish_column(b, r, gv, o, rb, fs, out, col, bins, base_at(cb + 4 * (int)tv.own_off4 + qbytes, col), [&](auto &&f) {
        for (int e = 0; e < m; e++) {
            int base, qual, score;
            if (fetch_ent(cb, ents[e], col, side, o, base, qual, score)) f(base, qual, score, e == fs.tmpl_k);
        }
    });
}

}  // namespace gcb
