// k_vote_tiled.cuh — the hot kernel: Pair::computeScore (pair.cpp:88-172) fused with
// Group::makeConsensus (group.cpp:320-579), sixteen template columns per lane.
//
// Work decomposition
//   CTA    = one tile: the consecutive clusters whose slab starts inside a window of the payload.  Two bulk
//            asynchronous copies (cp.async.bulk -> UBLKCP, one mbarrier) stage the tile's payload slab and
//            its VoteRead table in shared memory while the threads fetch the tile's FsDesc entries.
//   warp   = one bundle of family sides at a time (dynamic counter): L lanes per family side, L = the
//            tile's widest record in 16-column chunks, so three 150-base family sides share a warp.
//   lane   = sixteen consecutive columns of one family side: four 32-bit words of qualities and two words
//            of 4-bit bases per read, reduced with SIMD-in-word arithmetic; no per-base loop, no shuffles.
// A column is FAST when every voting read shows the template's base there, no read disagrees with its mate
// inside the pair overlap, and the best quality reaches moderateQuality.  With options for which that
// implies topScore >= baseScoreReq (`implied`, true for the reference's defaults) such a column is exactly
// group.cpp:421-427: second bin empty, only the quality (the maximum) is written.  Every other column is
// SLOW: it is queued, its reads are histogrammed into the sixteen bins of group.cpp:376-393 by eight
// threads per column (shared-memory atomics) and one thread per column then applies group.cpp:395-525
// through column_top()/column_arbitrate(), the same functions the generic kernel uses.
// Families whose voters all share the template's geometry (same length, no column shift, same overlap
// window: every fixed-length library) run a loop whose masks are hoisted out (FS_UNIFORM).
// Tiles that do not fit the tables (huge clusters, reads longer than 512, 16-bit field overflow) are handed
// to score_vote_kernel (k_score_vote.cuh) through ws.generic_tiles.
#pragma once

#include "k_score_vote.cuh"

namespace gcb {

constexpr int VT_THREADS = 256;
constexpr int VT_WARPS = VT_THREADS / WARP;
constexpr int VT_MAX_PAIRS = VT_THREADS;     // pair positions of a tile (one thread each in the prologue)
constexpr int VT_MAX_FS = 256;               // family sides of a tile
constexpr int VT_SLOW_CAP = 512;             // queued slow columns; more are decided inline by their owner
constexpr int VT_BIN_COLS = VT_THREADS / 8;  // slow columns histogrammed per pass (eight threads each)
constexpr int VT_CHUNK = 16;                 // columns per lane

// shared-memory map (bytes)
constexpr int VT_OFF_BAR = 0;
constexpr int VT_OFF_NSLOW = 8;
constexpr int VT_OFF_NOFIT = 12;
constexpr int VT_OFF_NFS = 16;
constexpr int VT_OFF_NEXT = 20;
constexpr int VT_OFF_LMAX = 24;
constexpr int VT_OFF_WSUM = 32;                                  // VT_WARPS warp totals of the compaction scan
constexpr int VT_OFF_ACC = 64;                                   // int32[VT_MAX_FS]
constexpr int VT_OFF_SLOW = VT_OFF_ACC + 4 * VT_MAX_FS;          // uint32[VT_SLOW_CAP]
constexpr int VT_OFF_BINS = VT_OFF_SLOW + 4 * VT_SLOW_CAP;       // int32[VT_BIN_COLS][16][4]
constexpr int VT_OFF_FT = VT_OFF_BINS + 4 * VT_BIN_COLS * 64;    // FsTile[VT_MAX_FS]
constexpr int VT_OFF_VR = VT_OFF_FT + 32 * VT_MAX_FS;            // VoteRead[2*VT_MAX_PAIRS]
constexpr int VT_OFF_SLAB = (VT_OFF_VR + 16 * 2 * VT_MAX_PAIRS + 127) & ~127;
constexpr int VT_SLAB_SLACK = 64;  // the hoisted loop reads whole words past a record's end (masked afterwards)
static_assert(VT_OFF_FT % 16 == 0 && VT_OFF_VR % 16 == 0 && VT_OFF_BINS % 16 == 0, "16-byte aligned tables");
constexpr int VT_MAX_SLAB = 200 * 1024;  // cbase4 / out4 are 16-bit counts of 4-byte units

// a family side as the tile sees it (shared memory only)
struct __align__(16) FsTile {
    uint16_t ent0;      // first VoteRead of this family side in the tile's table
    uint16_t m;
    uint16_t l_out;
    uint16_t len;
    uint16_t tmpl_k;
    uint8_t mode;
    uint8_t flags;
    uint16_t cbase4;    // the cluster's slab inside the staged tile, 4-byte units
    uint16_t out4;      // consensus record relative to the tile's first output byte, 4-byte units
    int64_t ref_nib0;   // nibble index of the template's pos in the packed genome (FS_REF_OK)
    int32_t slot;       // the family's group slot (result row)
    int32_t reserved;
};
static_assert(sizeof(FsTile) == 32 && sizeof(FsDesc) == 32 && sizeof(VoteRead) == 16, "table entry sizes");

// coverage counters of the CPU SIMT-check build (tests only): tiled tiles, generic tiles, voted columns, slow columns,
// uniform family sides, non-uniform family sides
#ifdef GCB_SIMT_CHECK
inline int64_t g_simt_counters[6] = {0, 0, 0, 0, 0, 0};
#define GCB_COUNT(k, n) (g_simt_counters[k] += (n))
#else
#define GCB_COUNT(k, n) ((void)0)
#endif

#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ void tile_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_copy(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#else
inline void tile_expect(uint64_t *, uint32_t) {}
inline void tile_copy(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
#endif

// ---- SIMD-in-word helpers ------------------------------------------------------------------------------
// Bases travel as big-endian nibble words: column k of an 8-column word sits in bits 28-4k..31-4k.
// Qualities stay little-endian: column k of a 4-column word is byte k.
// PTX prmt.b32 in its default mode: a selector nibble with bit 3 set replicates the SIGN of the selected byte
// (the __byte_perm intrinsic masks that bit off, so it cannot be used for the mask widening below)
#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
#else
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return ::simt::prmt(a, b, sel); }
#endif
// a 32-bit word of the CTA's shared memory at `addr` + IMM, where addr is a shared-window address (hot loop only:
// it keeps the address arithmetic in 32 bits and the immediate in the instruction)
#ifndef GCB_SIMT_CHECK
template <int IMM>
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ uint32_t smem_base(const void *smem) { return smem_u32(smem); }
#else
template <int IMM>
inline uint32_t lds32(uint32_t addr) { return *(const uint32_t *)(::simt::dyn_smem() + addr + IMM); }
inline uint32_t smem_base(const void *) { return 0u; }
#endif
GCB_DEV uint32_t bswap32(uint32_t w) { return __byte_perm(w, 0, 0x0123); }
GCB_DEV int clamp_int(int v, int lo, int hi) { return min(max(v, lo), hi); }
// columns >= s of an 8-column nibble word (s is clamped to 0..8)
GCB_DEV uint32_t nib_ge(int s) { return __funnelshift_rc(0xFFFFFFFFu, 0u, 4u * (unsigned)clamp_int(s, 0, 8)); }
// columns [a, z) of an 8-column nibble word
GCB_DEV uint32_t nib_range(int a, int z) { return nib_ge(a) & ~nib_ge(z); }
// a nibble mask (all-ones or all-zero nibbles) widened to the byte masks of its columns 0-3 and 4-7
GCB_DEV uint32_t bytes_lo(uint32_t nm) { return prmt(nm, nm << 4, 0xEAFBu); }
GCB_DEV uint32_t bytes_hi(uint32_t nm) { return prmt(nm, nm << 4, 0xC8D9u); }
// byte flags (0xFF / 0x00) of columns 0-3 (a) and 4-7 (b) narrowed to a nibble mask
GCB_DEV uint32_t nibs_of_bytes(uint32_t a, uint32_t b) {
    return (__byte_perm(a, b, 0x0246) & 0xF0F0F0F0u) | (__byte_perm(a, b, 0x1357) & 0x0F0F0F0Fu);
}
// one word of a record area of nw words, outside reads as 0
GCB_DEV uint32_t word_or_zero(const uint32_t *p, int w, int nw) { return (unsigned)w < (unsigned)nw ? p[w] : 0u; }

// sixteen qualities at read positions rp0..rp0+15 (any alignment, any sign); positions outside the area read as 0
GCB_DEV void fetch16q(const uint8_t *rec, int qbytes, int rp0, uint32_t q[4]) {
    const uint32_t *p = (const uint32_t *)rec;
    const int nw = qbytes >> 2, w0 = rp0 >> 2;
    const unsigned sh = (unsigned)(rp0 & 3) * 8u;
    uint32_t w[5];
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = word_or_zero(p, w0 + k, nw);
#pragma unroll
    for (int k = 0; k < 4; k++) q[k] = __funnelshift_r(w[k], w[k + 1], sh);
}
// sixteen base codes at read positions rp0..rp0+15 as two big-endian nibble words
GCB_DEV void fetch16b(const uint8_t *seq, int sbytes, int rp0, uint32_t &b0, uint32_t &b1) {
    const uint32_t *p = (const uint32_t *)seq;
    const int nw = sbytes >> 2, w0 = rp0 >> 3;
    const unsigned sh = (unsigned)(rp0 & 7) * 4u;
    const uint32_t a = bswap32(word_or_zero(p, w0, nw)), c = bswap32(word_or_zero(p, w0 + 1, nw)), d = bswap32(word_or_zero(p, w0 + 2, nw));
    b0 = __funnelshift_l(c, a, sh);
    b1 = __funnelshift_l(d, c, sh);
}

// base, rewritten quality and score of one read at template column i: pair.cpp:88-172 for one base,
// from the staged tables (the same function of the same bytes as fetch_base in k_score_vote.cuh)
GCB_DEV bool fetch_ent(const uint8_t *cb, const VoteRead &v, int i, int side, const gcb_options &o, int &base, int &qual, int &score) {
    if (v.own_off4 == VR_NO_VOTE || v.own_l == 0) return false;
    const int rp = i + v.shift;
    if (rp < 0 || rp >= v.own_l) return false;
    const uint8_t *q = cb + 4 * (int)v.own_off4;
    qual = q[rp];
    base = base_at(q + GCB_ALIGN4(v.own_l), rp);
    if (v.ov_len == VR_NO_OVERLAP_INFO) {
        score = sc8(o.score_moderate);
        return true;
    }
    const int k = rp - v.ov_own;
    if (k < 0 || k >= v.ov_len) {
        score = qual2score(o, qual);
        return true;
    }
    const int mp = v.ov_mate + k;
    if (mp < 0 || mp >= v.mate_l) {
        score = sc8(o.score_moderate);
        return true;
    }
    const uint8_t *mq = cb + 4 * (int)v.mate_off4;
    const int mqual = mq[mp];
    const int mbase = base_at(mq + GCB_ALIGN4(v.mate_l), mp);
    if (base == mbase) {
        score = sc8(qual2score(o, (qual + mqual) / 2) + 4);
    } else {
        const int lq = side == 0 ? qual : mqual, rq = side == 0 ? mqual : qual;
        const bool left_wins = lq >= rq;
        const bool mine = side == 0 ? left_wins : !left_wins;
        score = mine ? sc8(qual2score(o, lq >= rq ? lq - rq : rq - lq) - 3) : 0;
        qual = max(0, qual - mqual);
    }
    return true;
}

struct TileCtx {
    const BatchView *b;
    const ResultView *r;
    const GenomeView *gv;
    const gcb_options *o;
    const uint8_t *slab;
    const VoteRead *vr;
    const FsTile *ft;
    int32_t *acc;
    uint8_t *out0;  // out_payload + the tile's first output byte
};
GCB_DEV int fs_side(const FsTile &ft) { return (ft.flags & FS_SIDE1) ? 1 : 0; }

// group.cpp:376-393 for one read of one slow column: its vote goes into the column's sixteen bins
// {count, sum of scores, sum of qualities, best quality}
// Pair::qual2score (pair.cpp:77-86) as selects
GCB_DEV int qual2score_sel(const gcb_options &o, int q) {
    return q >= o.high_quality ? sc8(o.score_high) : q >= o.moderate_quality ? sc8(o.score_moderate) : q >= o.low_quality ? sc8(o.score_low) : sc8(o.score_bad);
}

// fetch_ent without divergent branches (the lanes of a warp histogram different reads of different columns, so every
// branch of fetch_ent would be walked by the whole warp): all cases are computed and selected.  Same results.
GCB_DEV bool fetch_vote(const uint8_t *cb, const VoteRead &v, int i, int side, const gcb_options &o, int &base, int &qual, int &score) {
    const int rp = i + v.shift;
    if (v.own_off4 == VR_NO_VOTE || rp < 0 || rp >= v.own_l) return false;
    const uint8_t *q = cb + 4 * (int)v.own_off4;
    const int ql = q[rp];
    base = base_at(q + GCB_ALIGN4(v.own_l), rp);
    const bool info = v.ov_len != VR_NO_OVERLAP_INFO;
    const int k = rp - v.ov_own, mp = v.ov_mate + k;
    const bool inwin = info && k >= 0 && k < v.ov_len;
    const bool mvalid = inwin && mp >= 0 && mp < v.mate_l;
    const int mpi = mvalid ? mp : 0;                                  // (any in-bounds byte when there is no mate base)
    const uint8_t *mq = cb + (mvalid ? 4 * (int)v.mate_off4 : 0);
    const int mql = mq[mpi];
    const int mbase = base_at(mq + (mvalid ? GCB_ALIGN4(v.mate_l) : 0), mpi);
    const bool match = base == mbase;
    const int lq = side == 0 ? ql : mql, rq = side == 0 ? mql : ql;
    const bool mine = side == 0 ? lq >= rq : !(lq >= rq);
    const int s_plain = qual2score_sel(o, ql);
    const int s_match = sc8(qual2score_sel(o, (ql + mql) / 2) + 4);
    const int s_mis = mine ? sc8(qual2score_sel(o, lq >= rq ? lq - rq : rq - lq) - 3) : 0;
    const int moderate = sc8(o.score_moderate);
    score = !info ? moderate : !inwin ? s_plain : !mvalid ? moderate : match ? s_match : s_mis;
    qual = (mvalid && !match) ? max(0, ql - mql) : ql;
    return true;
}

GCB_DEV void slow_histogram(const TileCtx &t, const FsTile &ft, int col, int e, int32_t *bins) {
    int base, qual, score;
    if (!fetch_vote(t.slab + 4 * (int)ft.cbase4, t.vr[ft.ent0 + e], col, fs_side(ft), *t.o, base, qual, score)) return;
    int32_t *bin = bins + 4 * base;
    atomicAdd(bin, 1);
    atomicAdd(bin + 1, score);
    atomicAdd(bin + 2, qual);
    atomicMax(bin + 3, qual);
}

// group.cpp:419-525 for one slow column once its top and second bins are known.  Thread-local.
// `acgt_maxq`: the best quality of codes 1, 2, 4, 8 in bytes 0..3 (0 when nobody showed the code).
GCB_DEV void slow_finish(const TileCtx &t, int f, int col, ColumnTop &top, int total, uint32_t acgt_maxq) {
    const gcb_options &o = *t.o;
    const FsTile ft = t.ft[f];
    const int side = fs_side(ft);
    const uint8_t *cb = t.slab + 4 * (int)ft.cbase4;
    const VoteRead *ents = t.vr + ft.ent0;
    const VoteRead tv = ents[ft.tmpl_k];
    const int qbytes = GCB_ALIGN4(ft.l_out);
    uint8_t *out = t.out0 + 4 * (int64_t)ft.out4;
    column_rules(o, top, total);
    int new_qual;
    if (top.fast) {
        new_qual = top.top.maxq;  // group.cpp:422-426: the base is NOT written
    } else {
        // the record's base before the vote: the template's own (pair.cpp rewrites qualities, never bases)
        const int obase = base_at(cb + 4 * (int)tv.own_off4 + qbytes, col);
        int ref4 = 0;
        if (ft.flags & FS_REF_OK) {  // group.cpp:430-439
            int refpos = col;
            if (!(ft.flags & FS_SIMPLE_CIGAR)) {
                const gcb_read_desc od = t.b->reads[t.r->groups[ft.slot].tmpl_read[side]];
                refpos = get_ref_offset(t.b->cigar + od.cigar_off, od.n_cigar, col);
            }
            const int64_t nib = ft.ref_nib0 + refpos;
            if (refpos >= 0 && nib >= 0 && (nib >> 1) < t.gv->packed_bytes) {  // the bound only guards malformed CIGARs
                const uint8_t two = t.gv->packed4[nib >> 1];
                ref4 = genome_nibble_to_bam((nib & 1) ? (two >> 4) : (two & 0xF));
            }
        }
        int rbq = 0;
        bool any_high = false;
        if (top.need_ref && ref4 != 0) {
            const int rmax = (int)((acgt_maxq >> (ref4 == 1 ? 0 : ref4 == 2 ? 8 : ref4 == 4 ? 16 : 24)) & 0xFFu);
            if (rmax >= 128) {  // `char refBaseQual` wraps: the scan order matters (group.cpp:474-490): template first
                int tb, tq, ts;
                if (fetch_ent(cb, tv, col, side, o, tb, tq, ts) && tb == ref4) {
                    if (tq > rbq) rbq = sc8(tq);
                    if (tq >= o.high_quality) any_high = true;
                }
                for (int e = 0; e < (int)ft.m; e++) {
                    int base, qual, score;
                    if (e == ft.tmpl_k || !fetch_ent(cb, ents[e], col, side, o, base, qual, score) || base != ref4) continue;
                    if (qual > rbq) rbq = sc8(qual);
                    if (qual >= o.high_quality) any_high = true;
                }
            } else {
                rbq = rmax;
                any_high = rmax >= o.high_quality;
            }
        }
        const ColumnOut co = column_arbitrate(o, top, ref4, rbq, any_high);
        if (obase != co.base) {  // group.cpp:509-524
            int d_mm = 0;
            if (ref4 != 0) {
                if (obase == ref4) d_mm = 1;
                else if (co.base == ref4) d_mm = -1;
            }
            atomicAdd(t.acc + f, 1 + d_mm * 65536);
            const int byte = col >> 1;
            const unsigned delta = ((unsigned)(obase ^ co.base) & 0xFu) << ((col & 1) ? 0 : 4);
            atomicXor((unsigned *)(out + qbytes + (byte & ~3)), delta << (8 * (byte & 3)));
        }
        new_qual = co.qual;
    }
    out[col] = (uint8_t)new_qual;
}

// beyond the voted columns the record keeps what it held (rewritten qualities)
GCB_DEV void slow_unvoted(const TileCtx &t, int f, int col) {
    const FsTile ft = t.ft[f];
    int obase = 0, oqual = 0, sc;
    fetch_ent(t.slab + 4 * (int)ft.cbase4, t.vr[ft.ent0 + ft.tmpl_k], col, fs_side(ft), *t.o, obase, oqual, sc);
    (t.out0 + 4 * (int64_t)ft.out4)[col] = (uint8_t)oqual;
}

// group.cpp:395-417 over complete bins, thread-local (queue overflow path)
GCB_DEV void slow_decide(const TileCtx &t, int f, int col, const int32_t *bins) {
    if (col >= (int)t.ft[f].len) {
        slow_unvoted(t, f, col);
        return;
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
    ColumnTop top = column_top(*t.o, obs, nobs, total);
    const uint32_t acgt = (uint32_t)(bins[4 * 1] > 0 ? bins[4 * 1 + 3] : 0) | ((uint32_t)(bins[4 * 2] > 0 ? bins[4 * 2 + 3] : 0) << 8) |
                          ((uint32_t)(bins[4 * 4] > 0 ? bins[4 * 4 + 3] : 0) << 16) | ((uint32_t)(bins[4 * 8] > 0 ? bins[4 * 8 + 3] : 0) << 24);
    slow_finish(t, f, col, top, total, acgt);
}

// The (score, quality sum, code) order of group.cpp:395-417 as one integer: the scans walk the sixteen bins,
// replace on a larger score or an equal score and a not-smaller quality sum, so they return the lexicographic
// maximum with ties going to the larger code; empty bins take part with (0, 0).
GCB_DEV unsigned long long bin_key(int score, int qual, int code) {
    return ((unsigned long long)(unsigned)(score + (1 << 23)) << 28) | ((unsigned long long)(unsigned)qual << 4) | (unsigned)code;
}
GCB_DEV unsigned long long max_u64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
GCB_DEV unsigned long long min_u64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

// One slow column by the eight lanes of an octet: group.cpp:376-393 (histogram, lanes stride the reads), then
// the two scans of group.cpp:395-417 as a top-2 reduction of the sixteen bin keys (two bins per lane, three
// shuffle steps), then lane 0 of the octet applies the rules.  Every lane of the warp must call it.
GCB_DEV void slow_octet(const TileCtx &t, bool active, int f, int col, int32_t *bins, int sub8) {
    FsTile ft;
    ft.m = 0; ft.len = 0; ft.ent0 = 0; ft.cbase4 = 0; ft.flags = 0;
    if (active) ft = t.ft[f];
    const bool voted = active && col < (int)ft.len;
    int4 *b4 = (int4 *)bins;
    const int4 zero = {0, 0, 0, 0};
    b4[2 * sub8] = zero;
    b4[2 * sub8 + 1] = zero;
    __syncwarp();
    if (voted)
        for (int e = sub8; e < (int)ft.m; e += 8) slow_histogram(t, ft, col, e, bins);
    __syncwarp();
    const int4 x0 = b4[2 * sub8], x1 = b4[2 * sub8 + 1];  // {count, sum of scores, sum of qualities, best quality}
    const unsigned long long k0 = bin_key(x0.y, x0.z, 2 * sub8), k1 = bin_key(x1.y, x1.z, 2 * sub8 + 1);
    unsigned long long top = max_u64(k0, k1), sec = min_u64(k0, k1);
    int total = x0.y + x1.y;
    for (int off = 1; off < 8; off <<= 1) {
        const unsigned long long ot = __shfl_xor_sync(FULL, top, off), os = __shfl_xor_sync(FULL, sec, off);
        total += __shfl_xor_sync(FULL, total, off);
        sec = max_u64(min_u64(top, ot), max_u64(sec, os));
        top = max_u64(top, ot);
    }
    if (!active || sub8 != 0) return;
    if (!voted) {
        slow_unvoted(t, f, col);
        return;
    }
    const int tb = (int)(top & 0xF), sb = (int)(sec & 0xF);
    const int4 T = b4[tb], S = b4[sb];
    ColumnTop ct;
    ct.top.base = tb; ct.top.cnt = T.x; ct.top.score = T.y; ct.top.qual = T.z; ct.top.maxq = T.w;
    ct.sec.base = sb; ct.sec.cnt = S.x; ct.sec.score = S.y; ct.sec.qual = S.z; ct.sec.maxq = S.w;
    const uint32_t acgt = (uint32_t)(b4[1].x > 0 ? b4[1].w : 0) | ((uint32_t)(b4[2].x > 0 ? b4[2].w : 0) << 8) |
                          ((uint32_t)(b4[4].x > 0 ? b4[4].w : 0) << 16) | ((uint32_t)(b4[8].x > 0 ? b4[8].w : 0) << 24);
    slow_finish(t, f, col, ct, total, acgt);
}

// a slow column decided by its owner alone (queue overflow): the histogram lives in local memory
GCB_DEV void slow_inline(const TileCtx &t, int f, int col) {
    int32_t bins[64];
    for (int k = 0; k < 64; k++) bins[k] = 0;
    const FsTile ft = t.ft[f];
    if (col < (int)ft.len)
        for (int e = 0; e < (int)ft.m; e++) {
            int base, qual, score;
            if (!fetch_ent(t.slab + 4 * (int)ft.cbase4, t.vr[ft.ent0 + e], col, fs_side(ft), *t.o, base, qual, score)) continue;
            bins[4 * base]++;
            bins[4 * base + 1] += score;
            bins[4 * base + 2] += qual;
            bins[4 * base + 3] = max(bins[4 * base + 3], qual);
        }
    slow_decide(t, f, col, bins);
}

// group.cpp:538-566: more than five new mismatches => the record keeps the template's bases and (rewritten) qualities
GCB_DEV void rollback_record(const TileCtx &t, int f) {
    const FsTile ft = t.ft[f];
    const uint8_t *cb = t.slab + 4 * (int)ft.cbase4;
    const VoteRead tv = t.vr[ft.ent0 + ft.tmpl_k];
    const int l_out = ft.l_out, qbytes = GCB_ALIGN4(l_out);
    uint8_t *out = t.out0 + 4 * (int64_t)ft.out4;
    const uint8_t *tseq = cb + 4 * (int)tv.own_off4 + qbytes;
    for (int col = 0; col < l_out; col++) {
        int base, qual, sc;
        fetch_ent(cb, tv, col, fs_side(ft), *t.o, base, qual, sc);
        out[col] = (uint8_t)qual;
    }
    for (int k = 0; k < (l_out + 1) >> 1; k++) out[qbytes + k] = tseq[k];
}

// what a lane needs to know about its sixteen columns of a record of l_out bases of which `len` are voted
struct ChunkMasks {
    uint32_t vn0, vn1;   // voted columns, nibble masks of columns 0-7 and 8-15
    uint32_t vb[4];      // voted columns, byte masks of the four quality words
    uint32_t rb[4];      // columns of the record (l_out), byte masks
    uint32_t kn0, kn1;   // nibbles the record keeps: its columns plus the odd tail nibble
    int nvote;
};
GCB_DEV ChunkMasks make_masks(int l_out, int len, int col0) {
    ChunkMasks c;
    c.nvote = clamp_int(len - col0, 0, VT_CHUNK);
    c.vn0 = nib_range(0, c.nvote);
    c.vn1 = nib_range(0, c.nvote - 8);
    c.vb[0] = bytes_lo(c.vn0); c.vb[1] = bytes_hi(c.vn0); c.vb[2] = bytes_lo(c.vn1); c.vb[3] = bytes_hi(c.vn1);
    const int nv = clamp_int(l_out - col0, 0, VT_CHUNK);
    const uint32_t rn0 = nib_range(0, nv), rn1 = nib_range(0, nv - 8);
    c.rb[0] = bytes_lo(rn0); c.rb[1] = bytes_hi(rn0); c.rb[2] = bytes_lo(rn1); c.rb[3] = bytes_hi(rn1);
    const int nk = clamp_int(2 * ((l_out + 1) >> 1) - col0, 0, VT_CHUNK);
    c.kn0 = nib_range(0, nk);
    c.kn1 = nib_range(0, nk - 8);
    return c;
}

__global__ void __launch_bounds__(VT_THREADS, 3) vote_tiled_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o,
                                                                int32_t slab_cap, int32_t implied) {
    GCB_DYN_SMEM(smem);
    uint64_t *bar = (uint64_t *)(smem + VT_OFF_BAR);
    int *s_nslow = (int *)(smem + VT_OFF_NSLOW);
    int *s_nofit = (int *)(smem + VT_OFF_NOFIT);
    int *s_nfs = (int *)(smem + VT_OFF_NFS);
    int *s_next = (int *)(smem + VT_OFF_NEXT);
    int *s_lmax = (int *)(smem + VT_OFF_LMAX);
    uint32_t *s_wsum = (uint32_t *)(smem + VT_OFF_WSUM);
    int32_t *s_acc = (int32_t *)(smem + VT_OFF_ACC);
    uint32_t *s_slow = (uint32_t *)(smem + VT_OFF_SLOW);
    int32_t *s_bins = (int32_t *)(smem + VT_OFF_BINS);
    FsTile *s_ft = (FsTile *)(smem + VT_OFF_FT);
    VoteRead *s_vr = (VoteRead *)(smem + VT_OFF_VR);
    uint8_t *slab = smem + VT_OFF_SLAB;
#define GCB_LDS32(off) (*(const uint32_t *)(smem + (off)))

    const int tid = (int)threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const TileDir t0 = ws.tile_dir[blockIdx.x], t1 = ws.tile_dir[blockIdx.x + 1];
    const int c0 = t0.c0, c1 = t1.c0;
    if (c0 >= c1) return;
    const int P0 = t0.p0, NP = t1.p0 - t0.p0;
    const int64_t slab_bytes = t1.slab0 - t0.slab0;
    if (NP > VT_MAX_PAIRS || slab_bytes > slab_cap) {  // not a tile for this kernel
        if (tid == 0) {
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = (int32_t)blockIdx.x;
            GCB_COUNT(1, 1);
        }
        return;
    }
    if (NP == 0) return;  // clusters without pairs emit nothing
    if (tid == 0) {
        tile_barrier_init(bar);
        *s_nslow = 0;
        *s_nofit = 0;
        *s_next = 0;
        *s_lmax = 1;
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t tb = 32u * (uint32_t)NP;
        tile_expect(bar, (uint32_t)slab_bytes + tb);
        if (slab_bytes > 0) tile_copy(slab, b.payload + t0.slab0, (uint32_t)slab_bytes, bar);
        tile_copy(s_vr, ws.vote_reads + 2 * (int64_t)P0, tb, bar);
    }
    // ---- prologue: one thread per pair position = per possible family slot; FsDesc is self-contained
    FsDesc fd[2];
    fd[0].mode = fd[1].mode = SIDE_NONE;
    fd[0].c = fd[1].c = c0;
    if (tid < NP) {  // slots that hold no family carry SIDE_NONE in side_mode and garbage in fs_desc
        const uint16_t modes = *(const uint16_t *)(ws.side_mode + 2 * (int64_t)(P0 + tid));
        fd[0] = ws.fs_desc[2 * (int64_t)(P0 + tid)];
        fd[1] = ws.fs_desc[2 * (int64_t)(P0 + tid) + 1];
        if ((modes & 0xFF) == SIDE_NONE) fd[0].mode = SIDE_NONE;
        if ((modes >> 8) == SIDE_NONE) fd[1].mode = SIDE_NONE;
    }
    const int64_t out_base0 = ws.scan_block[c0 / SCAN_BLOCK] + ws.cluster_out_off[c0];
    const bool live0 = fd[0].mode != SIDE_NONE, live1 = fd[1].mode != SIDE_NONE;
    int64_t c_slab = 0, c_out = 0;
    if (live0 || live1) {
        const int c = live0 ? fd[0].c : fd[1].c;
        c_slab = ws.slab_off[c] - t0.slab0;
        c_out = ws.scan_block[c / SCAN_BLOCK] + ws.cluster_out_off[c] - out_base0;
    }
    // compact index of the live family sides (exclusive scan of the live counts)
    int fidx0;
    {
        const uint32_t mine = (live0 ? 1u : 0u) + (live1 ? 1u : 0u);
        uint32_t incl = mine;
        for (int off = 1; off < WARP; off <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == WARP - 1) s_wsum[warp] = incl;
        __syncthreads();
        // exclusive scan of the warp totals by every warp for itself (VT_WARPS <= 32 lanes)
        const uint32_t wt = lane < VT_WARPS ? s_wsum[lane] : 0u;
        uint32_t wincl = wt;
        for (int off = 1; off < VT_WARPS; off <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, wincl, off);
            if (lane >= off) wincl += v;
        }
        const uint32_t pre = incl - mine + __shfl_sync(FULL, wincl - wt, warp);
        fidx0 = (int)pre;
        if (tid == VT_THREADS - 1) {
            *s_nfs = (int)(pre + mine);
            if (pre + mine > (uint32_t)VT_MAX_FS) *s_nofit = 1;
        }
    }
    if (live0 || live1) {
        int lneed = 1, fidx = fidx0;
        for (int side = 0; side < 2; side++) {
            if (fd[side].mode == SIDE_NONE) continue;
            const FsDesc d = fd[side];
            FsTile ft;
            ft.ent0 = (uint16_t)(2 * (d.mb - P0) + side * (int)d.m);
            ft.m = d.m;
            ft.l_out = d.l_out;
            ft.len = d.len;
            ft.tmpl_k = d.tmpl_k;
            ft.mode = d.mode;
            ft.flags = (uint8_t)(d.flags | (side ? FS_SIDE1 : 0));
            ft.cbase4 = (uint16_t)(c_slab >> 2);
            const int64_t orel = c_out + d.out_rel;
            ft.out4 = (uint16_t)(orel >> 2);
            ft.ref_nib0 = d.ref_nib0;
            ft.slot = P0 + tid;
            ft.reserved = 0;
            const int l = d.l_out;
            const int chunks = max((GCB_ALIGN4(l) + 15) >> 4, (GCB_ALIGN4((l + 1) >> 1) + 7) >> 3);
            if ((d.flags & FS_NOFIT) || (orel >> 2) > 0xFFFF || chunks > WARP) *s_nofit = 1;
            if (out_base0 + orel + record_bytes(l) > r.out_capacity) {
                raise_error(ws.error_flag, GCB_ERR_CAPACITY);
                ft.mode = SIDE_NONE;  // keeps its place in the table but is never voted
            }
            lneed = max(lneed, min(chunks, WARP));
            if (fidx < VT_MAX_FS) {
                s_ft[fidx] = ft;
                s_acc[fidx] = 0;
            }
            fidx++;
        }
        if (lneed > 1) atomicMax(s_lmax, lneed);
    }
    tile_wait(bar, 0);
    __syncthreads();
    if (*s_nofit) {  // the generic kernel takes the tile
        if (tid == 0) {
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = (int32_t)blockIdx.x;
            GCB_COUNT(1, 1);
        }
        return;
    }
    if (tid == 0) GCB_COUNT(0, 1);
    const int nfs = *s_nfs;

    TileCtx t;
    t.b = &b; t.r = &r; t.gv = &gv; t.o = &o;
    t.slab = slab; t.vr = s_vr; t.ft = s_ft; t.acc = s_acc;
    t.out0 = r.out_payload + out_base0;

    const uint32_t mod4 = 0x01010101u * (uint32_t)(o.moderate_quality & 0xFF);
    const uint32_t sbase = smem_base(smem);
    // (divisions of small numbers by multiply-and-shift: exact for numerators below 2^16 / divisor)
    const int L = *s_lmax;                              // lanes per family side, 1..32
    const int S = (int)((32u * ((65535u / (unsigned)L) + 1u)) >> 16);  // family sides per bundle = 32 / L
    const int nb = (int)(((unsigned)(nfs + S - 1) * ((65535u / (unsigned)S) + 1u)) >> 16);
    const int sub = (int)(((unsigned)lane * ((65535u / (unsigned)L) + 1u)) >> 16), j = lane - sub * L;
    const int col0 = VT_CHUNK * j;
    const int common_l = s_ft[0].l_out;  // the masks of the tile's usual record length are computed once
    const ChunkMasks cm_common = make_masks(common_l, common_l, col0);
    for (;;) {
        int bundle = 0;
        if (lane == 0) bundle = atomicAdd(s_next, 1);
        bundle = __shfl_sync(FULL, bundle, 0);
        if (bundle >= nb) break;
        const int f = bundle * S + sub;
        FsTile ft;
        ft.ent0 = 0; ft.m = 0; ft.l_out = 0; ft.len = 0; ft.tmpl_k = 0; ft.mode = SIDE_NONE; ft.flags = 0; ft.cbase4 = 0; ft.out4 = 0;
        if (sub < S && f < nfs) ft = s_ft[f];
        const int l_out = ft.l_out, len = ft.len;
        const int qbytes = GCB_ALIGN4(l_out), sbytes = GCB_ALIGN4((l_out + 1) >> 1);
        const bool mine = ft.mode != SIDE_NONE && col0 < max(qbytes, 2 * sbytes);  // this lane owns words of the record
        const int m = mine && ft.mode != SIDE_COPY ? (int)ft.m : 0;
        const int mmax = __reduce_max_sync(FULL, m);
        const int cb = VT_OFF_SLAB + 4 * (int)ft.cbase4;  // byte offsets into the CTA's shared memory
        const int ento = VT_OFF_VR + 16 * (int)ft.ent0;
        VoteRead tv = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t tbe0 = 0u, tbe1 = 0u;
        int trec = cb;
        if (mine) {
            tv = s_vr[ft.ent0 + ft.tmpl_k];
            trec = cb + 4 * (int)tv.own_off4;
            if (8 * j < sbytes) tbe0 = bswap32(GCB_LDS32(trec + qbytes + 8 * j));
            if (8 * j + 4 < sbytes) tbe1 = bswap32(GCB_LDS32(trec + qbytes + 8 * j + 4));
        }
        ChunkMasks cm = cm_common;
        if (l_out != common_l || len != l_out) cm = make_masks(l_out, len, col0);
        if (mine && j == 0 && ft.mode != SIDE_COPY) GCB_COUNT((ft.flags & FS_UNIFORM) ? 4 : 5, 1);
        // per-column maxima live in 16-bit lanes (VIMNMX.U16x2 is native, a per-byte maximum is seven instructions):
        // mo[k] tracks bytes 1 and 3 of quality word k in the high byte of each half, me[k] bytes 0 and 2 (word << 8)
        uint32_t mo[4] = {0u, 0u, 0u, 0u}, me[4] = {0u, 0u, 0u, 0u}, dis0 = 0u, dis1 = 0u;
        if (ft.flags & FS_UNIFORM) {
            // hoisted geometry: every voter is read at the template's columns and meets its mate at the same offset
            const int x = (int)tv.ov_own - col0;
            const int y = x - (int)tv.ov_mate;
            const int oa = max(max(0, x), y), oz = min(min(cm.nvote, x + (int)tv.ov_len), y + (int)tv.mate_l);
            const bool has_ov = tv.ov_len > 0 && oz > oa;
            const uint32_t om0 = has_ov ? nib_range(oa, oz) : 0u, om1 = has_ov ? nib_range(oa - 8, oz - 8) : 0u;
            const int mnw = GCB_ALIGN4((tv.mate_l + 1) >> 1) >> 2;
            const int ms = 0 - y, mw0 = ms >> 3;
            const unsigned msh = (unsigned)(ms & 7) * 4u;
            const bool p0 = has_ov && (unsigned)mw0 < (unsigned)mnw, p1 = has_ov && (unsigned)(mw0 + 1) < (unsigned)mnw,
                       p2 = has_ov && (unsigned)(mw0 + 2) < (unsigned)mnw;
            const uint32_t qbase = sbase + (uint32_t)(cb + col0), sdelta = (uint32_t)(qbytes - col0 + 8 * j),
                           mbase = sbase + (uint32_t)(cb + GCB_ALIGN4(tv.mate_l) + 4 * mw0);
            uint32_t ea = sbase + (uint32_t)ento;
            for (int e = 0; e < mmax; e++, ea += 16) {
                if (e >= m) continue;
                const uint32_t w = lds32<0>(ea);
                if ((w & 0xFFFFu) == VR_NO_VOTE) continue;
                const uint32_t qa = qbase + ((w & 0xFFFFu) << 2), sa = qa + sdelta;
                const uint32_t q0 = lds32<0>(qa), q1 = lds32<4>(qa), q2 = lds32<8>(qa), q3 = lds32<12>(qa);
                const uint32_t be0 = bswap32(lds32<0>(sa)), be1 = bswap32(lds32<4>(sa));
                mo[0] = __vmaxu2(mo[0], q0); me[0] = __vmaxu2(me[0], q0 << 8);
                mo[1] = __vmaxu2(mo[1], q1); me[1] = __vmaxu2(me[1], q1 << 8);
                mo[2] = __vmaxu2(mo[2], q2); me[2] = __vmaxu2(me[2], q2 << 8);
                mo[3] = __vmaxu2(mo[3], q3); me[3] = __vmaxu2(me[3], q3 << 8);
                dis0 |= be0 ^ tbe0;
                dis1 |= be1 ^ tbe1;
                if (has_ov) {  // pair.cpp:133-170: a base that differs from its mate's is never a fast column
                    const uint32_t ma = mbase + ((w >> 16) << 2);
                    const uint32_t a = p0 ? bswap32(lds32<0>(ma)) : 0u, c = p1 ? bswap32(lds32<4>(ma)) : 0u,
                                   d = p2 ? bswap32(lds32<8>(ma)) : 0u;
                    dis0 |= (be0 ^ __funnelshift_l(c, a, msh)) & om0;
                    dis1 |= (be1 ^ __funnelshift_l(d, c, msh)) & om1;
                }
            }
        } else {
            for (int e = 0; e < mmax; e++) {
                if (e >= m) continue;
                const VoteRead v = s_vr[ft.ent0 + e];
                if (v.own_off4 == VR_NO_VOTE || v.own_l == 0) continue;
                const int rp0 = col0 + v.shift;
                const int a = max(0, 0 - rp0), z = min(cm.nvote, (int)v.own_l - rp0);
                if (z <= a) continue;
                const uint8_t *rec = smem + cb + 4 * (int)v.own_off4;
                const int rq = GCB_ALIGN4(v.own_l);
                uint32_t q[4], be0, be1;
                fetch16q(rec, rq, rp0, q);
                fetch16b(rec + rq, GCB_ALIGN4((v.own_l + 1) >> 1), rp0, be0, be1);
                const uint32_t vm0 = nib_range(a, z), vm1 = nib_range(a - 8, z - 8);
                q[0] &= bytes_lo(vm0); q[1] &= bytes_hi(vm0); q[2] &= bytes_lo(vm1); q[3] &= bytes_hi(vm1);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    mo[k] = __vmaxu2(mo[k], q[k]);
                    me[k] = __vmaxu2(me[k], q[k] << 8);
                }
                dis0 |= (be0 ^ tbe0) & vm0;
                dis1 |= (be1 ^ tbe1) & vm1;
                if (v.ov_len > 0) {
                    // chunk column k pairs own index rp0+k with mate index k - y.  (Written with subtractions only:
                    // ptxas 12.9 dropped the negation when it folded max(a, max(x, -t)) into one VIMNMX3 on sm_100a —
                    // the PTX was right, the SASS and the B200 were not; profiles/r01_notes.md has the listing.)
                    const int x = (int)v.ov_own - rp0;  // first chunk column inside the overlap window
                    const int y = x - (int)v.ov_mate;   // first chunk column whose mate index is >= 0
                    const int oa = max(max(a, x), y);
                    const int oz = min(min(z, x + (int)v.ov_len), y + (int)v.mate_l);
                    if (oz > oa) {
                        const uint8_t *mrec = smem + cb + 4 * (int)v.mate_off4;
                        uint32_t mb0, mb1;
                        fetch16b(mrec + GCB_ALIGN4(v.mate_l), GCB_ALIGN4((v.mate_l + 1) >> 1), 0 - y, mb0, mb1);
                        dis0 |= (be0 ^ mb0) & nib_range(oa, oz);
                        dis1 |= (be1 ^ mb1) & nib_range(oa - 8, oz - 8);
                    }
                }
            }
        }
        if (!mine) continue;
        // ---- what the record gets: qualities = the maxima (fast columns), bases = the template's
        uint32_t oq[4];
        uint32_t slow0 = 0u, slow1 = 0u;
        if (ft.mode == SIDE_COPY) {  // group.cpp:73-77: the record itself
#pragma unroll
            for (int k = 0; k < 4; k++) oq[k] = col0 + 4 * k < qbytes ? GCB_LDS32(trec + col0 + 4 * k) : 0u;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) oq[k] = prmt(mo[k], me[k], 0x3715u) & cm.vb[k];  // (the hoisted loop read whole words)
            dis0 &= cm.vn0;
            dis1 &= cm.vn1;
            GCB_COUNT(2, cm.nvote);
            if (implied && len == l_out) {
                const uint32_t lowq0 = nibs_of_bytes(~__vcmpgeu4(oq[0], mod4), ~__vcmpgeu4(oq[1], mod4));
                const uint32_t lowq1 = nibs_of_bytes(~__vcmpgeu4(oq[2], mod4), ~__vcmpgeu4(oq[3], mod4));
                slow0 = (dis0 | lowq0) & cm.vn0;
                slow1 = (dis1 | lowq1) & cm.vn1;
            } else {  // without `implied`, or with columns that are not voted, every column of the record is slow
                slow0 = nibs_of_bytes(cm.rb[0], cm.rb[1]);
                slow1 = nibs_of_bytes(cm.rb[2], cm.rb[3]);
            }
        }
        uint8_t *out = t.out0 + 4 * (int64_t)ft.out4;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (col0 + 4 * k < qbytes) *(uint32_t *)(out + col0 + 4 * k) = oq[k] & cm.rb[k];
        if (8 * j < sbytes) *(uint32_t *)(out + qbytes + 8 * j) = bswap32(tbe0 & cm.kn0);
        if (8 * j + 4 < sbytes) *(uint32_t *)(out + qbytes + 8 * j + 4) = bswap32(tbe1 & cm.kn1);
        // queue the slow columns (one or two per family side of a clean library)
        for (int wsel = 0; wsel < 2; wsel++) {
            uint32_t sm = wsel ? slow1 : slow0;
            while (sm != 0u) {
                const int k = __clz((int)sm) >> 2;
                sm &= ~(0xF0000000u >> (4 * k));
                const int col = col0 + 8 * wsel + k;
                GCB_COUNT(3, 1);
                const int idx = atomicAdd(s_nslow, 1);
                if (idx < VT_SLOW_CAP) s_slow[idx] = ((uint32_t)f << 16) | (uint32_t)col;
                else slow_inline(t, f, col);  // queue full: this lane owns the chunk's words
            }
        }
    }
    __syncthreads();
    // ---- slow columns: one octet (eight lanes) per column, thirty-two columns per pass; an octet owns its bins
    {
        const int n = min(*s_nslow, VT_SLOW_CAP);
        const int ci = tid >> 3, sub8 = tid & 7;
        for (int base = 0; base < n; base += VT_BIN_COLS) {
            const bool active = base + ci < n;
            const uint32_t code = active ? s_slow[base + ci] : 0u;
            slow_octet(t, active, (int)(code >> 16), (int)(code & 0xFFFFu), s_bins + 64 * ci, sub8);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- per family side: diff, mismatchInc, rollback, absolute output offset
    if (live0 || live1) {
        int fidx = fidx0;
        for (int side = 0; side < 2; side++) {
            if (fd[side].mode == SIDE_NONE) continue;
            const int f = fidx++;
            const FsTile ft = s_ft[f];
            if (ft.mode == SIDE_NONE) continue;
            const int acc = s_acc[f];
            const int diff = acc & 0xFFFF, mm = (acc - diff) >> 16;
            if (mm > 5) rollback_record(t, f);
            gcb_group_result *gr = r.groups + (P0 + tid);
            gr->diff[side] = diff;
            gr->mismatch_inc[side] = mm;
            gr->out_off[side] = out_base0 + 4 * (int64_t)ft.out4;
        }
    }
#undef GCB_LDS32
}

}  // namespace gcb
