// k_slow_columns.cuh — the slow columns of the vote, decided at full occupancy: group.cpp:376-525 for every column the
// ring kernel (k_vote_ring.cuh) could not finish in the word.
//
// The ring kernel keeps a tile's payload in shared memory only as long as its warps vote; what a slow column needs of the
// tile — per read of the family side its quality, its base nibble, its mate's quality and base nibble and where
// pair.cpp:121-170 puts the column (no overlap information / outside the overlap / mate base present / mate index out of
// range): 4 bytes per read behind a 32-byte self-contained header — is extracted by the warp that closes the tile, one
// thread per column, into a global queue (one 64-bit atomic per tile reserves the tile's records and words in one
// counter).  slow_columns_kernel then takes one record per thread: score per read (pair.cpp), three-bin
// register histogram (group.cpp:376-393), top-2 selection (group.cpp:395-417), the rules and the reference arbitration
// (group.cpp:419-525); it patches the consensus record and adds to the family side's diff / mismatchInc (atomics on the
// result row; the ring kernel zeroed them).  The deciding is a chain of dependent small loads: it wants many resident
// warps, which the one-CTA-per-SM ring cannot give it.
#pragma once

#include "vote_tile.cuh"

namespace gcb {

constexpr int VQ_NQ = 1;                       // slow-column queues; the tiles of CTA b use queue b % VQ_NQ (one reservation per tile: one
                                               // counter takes them all, and the whole capacity is there for whoever needs it)
constexpr uint32_t VQ_INVALID = 0xFFFFFFFFu;   // index entry of a reservation that did not fit
constexpr int VQ_SLOW_THREADS = 128;
constexpr int VQ_SLOW_CTAS = 148 * 8;          // slow_columns_kernel strides over the records

struct SlowQueues {
    unsigned long long *count;   // [VQ_NQ] records << 32 | words reserved so far (may run past the capacity)
    uint32_t *words;             // [VQ_NQ][cap_words] records (see SR_HDR_WORDS)
    uint32_t *index;             // [VQ_NQ][cap_recs] word offset of every record inside its queue, VQ_INVALID = none
    uint32_t cap_words, cap_recs;
};

// record: SR_HDR_WORDS header words, then n entries (one per read of the family side), padded to a multiple of 4 words.
// The header is self-contained (slow_columns_kernel needs no table lookup):
//   [0] 2 * slot + side   [1] col | n << 16   [2] tmpl_k | flags << 16   [3] l_out
//   [4..5] absolute offset of the consensus record in out_payload   [6..7] FsTile.ref_nib0
constexpr int SR_HDR_WORDS = 8;
constexpr uint32_t SR_UNVOTED = 1u;        // column beyond the voted length: the record keeps the template's (rewritten) quality
constexpr uint32_t SR_REF_OK = 2u;         // FS_REF_OK
constexpr uint32_t SR_SIMPLE_CIGAR = 4u;   // FS_SIMPLE_CIGAR
// entry: quality | mate quality << 8 | base << 16 | mate base << 20 | state << 24 | SE_VOTES
constexpr uint32_t SE_VOTES = 1u << 26;
constexpr uint32_t SE_NO_INFO = 0u, SE_PLAIN = 1u, SE_MATE = 2u, SE_NO_MATE_BASE = 3u;

GCB_DEV uint32_t slow_rec_words(int m) { return (uint32_t)SR_HDR_WORDS + (((uint32_t)m + 3u) & ~3u); }
GCB_DEV void slow_write_header(uint32_t *rec, const FsTile &ft, int col, int64_t out_abs) {
    const uint32_t side = (ft.flags & FS_SIDE1) ? 1u : 0u;
    const uint32_t fl = (col >= (int)ft.len ? SR_UNVOTED : 0u) | ((ft.flags & FS_REF_OK) ? SR_REF_OK : 0u) |
                        ((ft.flags & FS_SIMPLE_CIGAR) ? SR_SIMPLE_CIGAR : 0u);
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
    const int lq = side == 0 ? ql : mql, rq = side == 0 ? mql : ql;
    const bool mine = side == 0 ? lq >= rq : !(lq >= rq);
    const int s_match = sc8(t.q2s((ql + mql) / 2) + 4);                        // pair.cpp:147-152
    const int s_mis = mine ? sc8(t.q2s(lq >= rq ? lq - rq : rq - lq) - 3) : 0;  // pair.cpp:153-169
    const bool mism = st == SE_MATE && base != mbase;
    score = st == SE_MATE ? (mism ? s_mis : s_match) : st == SE_PLAIN ? t.q2s(ql) : t.sm;
    qual = mism ? max(0, ql - mql) : ql;
    return true;
}

// ------------------------------------------------------------------------------------------------
// A fourth distinct code in one column: the sixteen-bin histogram in local memory (group.cpp:376-417 as written).
__device__ __noinline__ void slow_record_wide(const gcb_options &o, const uint32_t *ents, int n, int side, ColumnTop &ct, int &total_out,
                                              uint32_t &acgt_out) {
    const ScoreTab tab(o);
    int32_t bins[64];
    for (int k = 0; k < 64; k++) bins[k] = 0;
    for (int e = 0; e < n; e++) {
        int base, qual, score;
        if (!slow_decode(tab, ents[e], side, base, qual, score)) continue;
        bins[4 * base]++;
        bins[4 * base + 1] += score;
        bins[4 * base + 2] += qual;
        bins[4 * base + 3] = max(bins[4 * base + 3], qual);
    }
    VoteBin obs[16];
    int nobs = 0, total = 0;
    for (int k = 0; k < 16; k++) {
        const int cnt = bins[4 * k];
        if (cnt > 0) {
            obs[nobs].base = k; obs[nobs].cnt = cnt; obs[nobs].score = bins[4 * k + 1]; obs[nobs].qual = bins[4 * k + 2]; obs[nobs].maxq = bins[4 * k + 3];
            total += obs[nobs].score;
            nobs++;
        }
    }
    ct = column_top(o, obs, nobs, total);
    total_out = total;
    acgt_out = (uint32_t)(bins[4 * 1] > 0 ? bins[4 * 1 + 3] : 0) | ((uint32_t)(bins[4 * 2] > 0 ? bins[4 * 2 + 3] : 0) << 8) |
               ((uint32_t)(bins[4 * 4] > 0 ? bins[4 * 4 + 3] : 0) << 16) | ((uint32_t)(bins[4 * 8] > 0 ? bins[4 * 8 + 3] : 0) << 24);
}

// group.cpp:376-525 for one queued column
GCB_DEV void slow_record(const BatchView &b, const ResultView &r, const GenomeView &gv, const gcb_options &o, const ScoreTab &tab,
                         const RollbackList &rb, const uint32_t *rec) {
    const uint4 ha = ((const uint4 *)rec)[0], hc = ((const uint4 *)rec)[1];
    const uint32_t fsid = ha.x, w1 = ha.y, w2 = ha.z;
    const int col = (int)(w1 & 0xFFFFu), n = (int)(w1 >> 16), tmpl_k = (int)(w2 & 0xFFFFu);
    const uint32_t flags = w2 >> 16;
    const uint32_t *ents = rec + SR_HDR_WORDS;
    const int side = (int)(fsid & 1u);
    const int qbytes = GCB_ALIGN4((int)ha.w);
    uint8_t *out = r.out_payload + (int64_t)(((uint64_t)hc.y << 32) | hc.x);
    const int64_t ref_nib0 = (int64_t)(((uint64_t)hc.w << 32) | hc.z);
    GCB_COUNT(3, 1);
    if (flags & SR_UNVOTED) {  // beyond the voted columns the record keeps what it held (rewritten qualities)
        int obase = 0, oqual = 0, sc;
        slow_decode(tab, ents[tmpl_k], side, obase, oqual, sc);
        out[col] = (uint8_t)oqual;
        return;
    }
    Bins3 bins;
    bins.init();
    for (int e = 0; e < n; e += 4) {  // (records are padded to whole 16-byte groups of entries)
        const uint4 v = *(const uint4 *)(ents + e);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int base, qual, score;
            if (e + k < n && slow_decode(tab, w[k], side, base, qual, score)) bins.add(base, qual, score);
        }
    }
    ColumnTop ct;
    int total = bins.total;
    uint32_t acgt = 0;
    if (bins.overflow) {
        slow_record_wide(o, ents, n, side, ct, total, acgt);
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
        // the record's base before the vote: the template's own (pair.cpp rewrites qualities, never bases)
        const int obase = (int)((ents[tmpl_k] >> 16) & 0xFu);
        int ref4 = 0;
        if (flags & SR_REF_OK) {  // group.cpp:430-439
            int refpos = col;
            if (!(flags & SR_SIMPLE_CIGAR)) {
                const gcb_read_desc od = b.reads[r.groups[fsid >> 1].tmpl_read[side]];
                refpos = get_ref_offset(b.cigar + od.cigar_off, od.n_cigar, col);
            }
            const int64_t nib = ref_nib0 + refpos;
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
                int tb, tq, ts;
                if (slow_decode(tab, ents[tmpl_k], side, tb, tq, ts) && tb == ref4) {
                    if (tq > rbq) rbq = sc8(tq);
                    if (tq >= o.high_quality) any_high = true;
                }
                for (int e = 0; e < n; e++) {
                    int base, qual, score;
                    if (e == tmpl_k || !slow_decode(tab, ents[e], side, base, qual, score) || base != ref4) continue;
                    if (qual > rbq) rbq = sc8(qual);
                    if (qual >= o.high_quality) any_high = true;
                }
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
            gcb_group_result *gr = r.groups + (fsid >> 1);  // (fsid = 2 * slot + side; the ring kernel zeroed both counters)
            atomicAdd(&gr->diff[side], 1);
            if (d_mm != 0) {
                const int before = atomicAdd(&gr->mismatch_inc[side], d_mm);
                if (d_mm > 0 && before == 5) {  // more than five new mismatches so far: vote_rollback_kernel looks at the final count
                    const int k = atomicAdd(rb.count, 1);
                    if (k < rb.cap) rb.list[k] = (int32_t)fsid;
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

__global__ void __launch_bounds__(VQ_SLOW_THREADS) slow_columns_kernel(BatchView b, ResultView r, GenomeView gv, gcb_options o, SlowQueues sq,
                                                                       RollbackList rb) {
    // the records of all queues as one index space, so that every warp but the last is full
    __shared__ uint32_t s_first[VQ_NQ + 1];
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int q = 0; q < VQ_NQ; q++) {
            s_first[q] = run;
            const uint32_t reserved = (uint32_t)(sq.count[q] >> 32);
            run += reserved < sq.cap_recs ? reserved : sq.cap_recs;
        }
        s_first[VQ_NQ] = run;
    }
    __syncthreads();
    const uint32_t total = s_first[VQ_NQ];
    const ScoreTab tab(o);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int q = 0;
        while (q + 1 < VQ_NQ && s_first[q + 1] <= i) q++;
        const uint32_t off = sq.index[(size_t)q * sq.cap_recs + (i - s_first[q])];
        if (off == VQ_INVALID) continue;
        slow_record(b, r, gv, o, tab, rb, sq.words + (size_t)q * sq.cap_words + off);
    }
}

}  // namespace gcb
