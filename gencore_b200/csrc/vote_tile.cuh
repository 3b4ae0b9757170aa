// vote_tile.cuh — what the tile-based vote (k_vote_ring.cuh) is built from: the per-tile tables, the SIMD-in-word
// helpers of the sixteen-columns-per-lane fast path, Pair::computeScore for one base (pair.cpp:88-172), the
// mbarrier / bulk-copy wrappers, the once-per-batch tile preparation kernel and the > 5-mismatch rollback kernel
// (group.cpp:538-566).
//
// A TILE is the run of consecutive clusters whose payload slab starts inside one window of the payload
// (umi_group_kernel writes the directory).  tile_prep2_kernel turns the directory into a 48-byte header per tile
// plus a compact list of the tile's live family sides (FsTile); the vote kernel stages the header's slab, the
// tile's VoteRead table and that list into shared memory with three bulk copies and never looks at a cluster again.
#pragma once

#include "k_score_vote.cuh"

namespace gcb {

constexpr int VT_CHUNK = 16;       // columns per lane
constexpr int VT_SLAB_SLACK = 64;  // the hoisted loop reads whole words past a record's end (masked afterwards)
constexpr int VT_MAX_SLAB = 200 * 1024;  // cbase4 / out4 / own_off4 are 16-bit counts of 4-byte units
constexpr int VS_MAX_PAIRS = 1024;  // pair positions of a tile (its VoteRead table has two entries per position)
constexpr int VS_MAX_FS = 192;      // family sides of a tile (eight bits in a slow-column list entry)
constexpr int VS_PREP_THREADS = 128;  // tile_prep2_kernel: four tiles per CTA, one warp each

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

struct __align__(16) TileHdr2 {
    int64_t out_base0;   // first output byte of the tile
    int64_t slab0;       // payload offset of the tile's slab
    int32_t slab_bytes;
    int32_t p0, np;      // pair positions
    int32_t nfs;         // live family sides (FsTile entries at fs_tiles[2*p0 ..]); 0 = nothing for the vote kernel
    int32_t lanes;       // lanes per family side: the tile's widest record in 16-column chunks
    int32_t common_l;    // l_out of the first family side (mask set computed once per tile)
    int32_t per_bundle;  // family sides per warp pass = 32 / lanes
    int32_t n_bundles;   // ceil(nfs / per_bundle)
};
static_assert(sizeof(TileHdr2) == 48, "tile header size");

// the family sides whose mismatchInc went from 5 to 6 while their slow columns were decided: rollback candidates
struct RollbackList {
    int32_t *list;   // [cap] 2 * slot + side
    int32_t *count;  // [1] entries appended (may run past cap: then every family side of the view is checked)
    int32_t cap;
};

// ---- mbarrier / bulk-copy wrappers.  Under SIMT-check (one OS thread, fibers) a barrier is a count of completed
// phases and a wait yields to the other fibers until the phase it names is over; copies are synchronous.
#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ void tile_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_copy(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pipe_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pipe_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pipe_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pipe_commit(uint64_t *) {}  // the hardware completes the phase when the bytes have landed
// one poll; the hardware may suspend the thread for up to `ns` nanoseconds before it answers
__device__ __forceinline__ bool pipe_try_wait(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void pipe_progress() {}
__device__ __forceinline__ void pipe_relax(uint32_t ns) { __nanosleep(ns); }  // between polls of a spin wait
#else
inline void tile_expect(uint64_t *, uint32_t) {}
inline void tile_copy(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
// bits 0..15 completed phases, 16..31 arrivals of the current phase, 32..47 arrivals a phase needs
inline void pipe_init(uint64_t *bar, int count) { *bar = (uint64_t)count << 32; }
inline void pipe_fence_init() {}
inline void pipe_arrive(uint64_t *bar) {
    uint64_t v = *bar;
    const uint64_t need = (v >> 32) & 0xFFFF;
    uint64_t arrived = ((v >> 16) & 0xFFFF) + 1, done = v & 0xFFFF;
    if (arrived == need) {
        arrived = 0;
        done = (done + 1) & 0xFFFF;
    }
    *bar = (need << 32) | (arrived << 16) | done;
    ::simt::st().progress++;
}
inline void pipe_commit(uint64_t *bar) { pipe_arrive(bar); }  // the copies above were synchronous
inline bool pipe_try_wait(uint64_t *bar, uint32_t parity, uint32_t) { return ((*bar) & 1u) != parity; }
inline void pipe_progress() { ::simt::st().progress++; }
inline void pipe_relax(uint32_t) { ::simt::relax(); }
#endif
GCB_DEV void pipe_expect(uint64_t *bar, uint32_t bytes) { tile_expect(bar, bytes); }
GCB_DEV void pipe_wait(uint64_t *bar, uint32_t parity, uint32_t ns) {
    while (!pipe_try_wait(bar, parity, ns)) pipe_relax(20u);
}

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
// a 32-bit word of the CTA's shared memory at `addr`, a shared-window address (hot loop only: the address arithmetic
// stays in 32 bits)
__device__ __forceinline__ uint32_t lds32r(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ uint32_t smem_base(const void *smem) { return smem_u32(smem); }
#else
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return ::simt::prmt(a, b, sel); }
inline uint32_t lds32r(uint32_t addr) { return *(const uint32_t *)(::simt::dyn_smem() + addr); }
template <int IMM>
inline uint32_t lds32(uint32_t addr) { return *(const uint32_t *)(::simt::dyn_smem() + addr + IMM); }
template <int IMM>
inline uint32_t lds16(uint32_t addr) { return *(const uint16_t *)(::simt::dyn_smem() + addr + IMM); }
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
// byte flags (bit 7 of every byte) of columns 0-3 (a) and 4-7 (b) widened to a big-endian nibble mask
GCB_DEV uint32_t nibs_of_flags(uint32_t a, uint32_t b) {
    return (prmt(a, b, 0x8ACEu) & 0xF0F0F0F0u) | (prmt(a, b, 0x9BDFu) & 0x0F0F0F0Fu);
}
// bit 7 of every byte of x that is >= the byte of t4 (t4 = four copies of a threshold <= 128)
GCB_DEV uint32_t bytes_ge_flags(uint32_t x, uint32_t t4) { return (((x | 0x80808080u) - t4) | x) & 0x80808080u; }
// the eight column flags of a nibble word (bit 28 - 4k = column k) gathered into bits 7 - k
GCB_DEV uint32_t nib_flags_to_byte(uint32_t x) {
    x &= 0x11111111u;
    x = (x | (x >> 3)) & 0x03030303u;
    x = (x | (x >> 6)) & 0x000F000Fu;
    return (x | (x >> 12)) & 0xFFu;
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

// ---- Pair::computeScore for one base --------------------------------------------------------------------
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

GCB_DEV int fs_side(const FsTile &ft) { return (ft.flags & FS_SIDE1) ? 1 : 0; }

// Pair::qual2score (pair.cpp:77-86) with the thresholds and the four scores in registers
struct ScoreTab {
    int hq, mq, lq, sh, sm, sl, sb;
    GCB_DEV explicit ScoreTab(const gcb_options &o)
        : hq(o.high_quality), mq(o.moderate_quality), lq(o.low_quality), sh(sc8(o.score_high)), sm(sc8(o.score_moderate)), sl(sc8(o.score_low)),
          sb(sc8(o.score_bad)) {}
    GCB_DEV int q2s(int q) const { return q >= hq ? sh : q >= mq ? sm : q >= lq ? sl : sb; }
};

// fetch_ent without divergent branches (the lanes of a warp histogram different reads of different columns, so every
// branch of fetch_ent would be walked by the whole warp): all cases are computed and selected.  Same results.
GCB_DEV bool fetch_vote(const uint8_t *cb, const VoteRead &v, int i, int side, const ScoreTab &t, int &base, int &qual, int &score) {
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
    const int s_plain = t.q2s(ql);
    const int s_match = sc8(t.q2s((ql + mql) / 2) + 4);
    const int s_mis = mine ? sc8(t.q2s(lq >= rq ? lq - rq : rq - lq) - 3) : 0;
    score = !info ? t.sm : !inwin ? s_plain : !mvalid ? t.sm : match ? s_match : s_mis;
    qual = (mvalid && !match) ? max(0, ql - mql) : ql;
    return true;
}

// The (score, quality sum, code) order of group.cpp:395-417 as one integer: the scans walk the sixteen bins,
// replace on a larger score or an equal score and a not-smaller quality sum, so they return the lexicographic
// maximum with ties going to the larger code; empty bins take part with (0, 0).
GCB_DEV unsigned long long bin_key(int score, int qual, int code) {
    return ((unsigned long long)(unsigned)(score + (1 << 23)) << 28) | ((unsigned long long)(unsigned)qual << 4) | (unsigned)code;
}
GCB_DEV unsigned long long max_u64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
GCB_DEV unsigned long long min_u64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

// group.cpp:376-393 for up to three distinct codes, in registers and without branches (a fourth code raises `overflow`)
struct Bins3 {
    int b0, b1, b2;              // codes (-1 = free)
    int c0, c1, c2;              // counts
    int s0, s1, s2;              // score sums
    int q0, q1, q2;              // quality sums
    int x0, x1, x2;              // best qualities
    int total;
    bool overflow;
    GCB_DEV void init() {
        b0 = b1 = b2 = -1;
        c0 = c1 = c2 = s0 = s1 = s2 = q0 = q1 = q2 = x0 = x1 = x2 = 0;
        total = 0;
        overflow = false;
    }
    // bin 0 may be SEEDED with a code nobody has shown yet (the template's base: the code most reads of a column show); an empty
    // bin takes part in the top-2 selection exactly as the code would without a bin — with (0, 0, code)
    GCB_DEV void seed(int base) { b0 = base; }
    GCB_DEV void add(int base, int qual, int score) {
        total += score;
        if (base == b0) {  // the usual read: five instructions instead of forty (a branch: the threads of a warp mostly agree)
            c0++; s0 += score; q0 += qual; x0 = max(x0, qual);
            return;
        }
        if (base == b1) {  // (a slow column usually shows two codes: the general update is left with the first sight of a code)
            c1++; s1 += score; q1 += qual; x1 = max(x1, qual);
            return;
        }
        // the bin that holds the code, else the first free one
        const bool h0 = false, h1 = false, h2 = b2 == base;
        const bool hit = h0 || h1 || h2;
        const bool u0 = h0 || (!hit && b0 < 0);
        const bool u1 = h1 || (!hit && b0 >= 0 && b1 < 0);
        const bool u2 = h2 || (!hit && b0 >= 0 && b1 >= 0 && b2 < 0);
        overflow = overflow || !(u0 || u1 || u2);
        // (selects, not branches: the threads of a warp histogram different columns)
        b0 = u0 ? base : b0; c0 += u0 ? 1 : 0; s0 += u0 ? score : 0; q0 += u0 ? qual : 0; x0 = max(x0, u0 ? qual : 0);
        b1 = u1 ? base : b1; c1 += u1 ? 1 : 0; s1 += u1 ? score : 0; q1 += u1 ? qual : 0; x1 = max(x1, u1 ? qual : 0);
        b2 = u2 ? base : b2; c2 += u2 ? 1 : 0; s2 += u2 ? score : 0; q2 += u2 ? qual : 0; x2 = max(x2, u2 ? qual : 0);
    }
    GCB_DEV VoteBin bin(int k) const {
        VoteBin v;
        v.base = k == 0 ? b0 : k == 1 ? b1 : b2;
        v.cnt = k == 0 ? c0 : k == 1 ? c1 : c2;
        v.score = k == 0 ? s0 : k == 1 ? s1 : s2;
        v.qual = k == 0 ? q0 : k == 1 ? q1 : q2;
        v.maxq = k == 0 ? x0 : k == 1 ? x1 : x2;
        return v;
    }
};

// the same for up to two distinct codes (a third raises `overflow`): what most slow columns need
struct Bins2 {
    int bA, bB, cA, cB, sA, sB, qA, qB, xA, xB, total;
    bool overflow;
    GCB_DEV void init() {
        bA = bB = -1;
        cA = cB = sA = sB = qA = qB = xA = xB = total = 0;
        overflow = false;
    }
    GCB_DEV void seed(int base) { bA = base; }  // (see Bins3::seed)
    GCB_DEV void add(int base, int qual, int score) {
        total += score;
        if (base == bA) {
            cA++; sA += score; qA += qual; xA = max(xA, qual);
            return;
        }
        if (base == bB) {
            cB++; sB += score; qB += qual; xB = max(xB, qual);
            return;
        }
        const bool uA = bA < 0;
        const bool uB = !uA && (bB < 0 || bB == base);
        overflow = overflow || !(uA || uB);
        bA = uA ? base : bA; cA += uA ? 1 : 0; sA += uA ? score : 0; qA += uA ? qual : 0; xA = max(xA, uA ? qual : 0);
        bB = uB ? base : bB; cB += uB ? 1 : 0; sB += uB ? score : 0; qB += uB ? qual : 0; xB = max(xB, uB ? qual : 0);
    }
};

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

// ------------------------------------------------------------------------------------------------
// One WARP per tile, 32 pair positions per pass: the tile's bookkeeping, once per batch.  The live family sides are
// compacted with ballots (no shared memory, no CTA barrier); many tiles per SM keep their chains of dependent loads
// (directory -> side modes -> family-side descriptors -> cluster offsets) in flight at once.  `max_need`: the largest
// shared-memory allocation a tile of this view takes in the vote kernel.
#ifndef GCB_VT_DEEP
#define GCB_VT_DEEP 24
#endif
GCB_HD bool tile_is_deep(int32_t nfs, int32_t np) { return nfs > 0 && 2 * np >= GCB_VT_DEEP * nfs; }  // 24 pairs or more per family side on average
GCB_HD int32_t tile_smem_need(int32_t nfs, int32_t np, int32_t slab_bytes, int32_t lanes) {
    // family-side list, VoteRead table, slab + slack; a deep tile also its slow-column list and the list's prefix sums (one entry
    // per family side and lane)
    return ((32 * nfs + 127) & ~127) + ((32 * np + 127) & ~127) + ((slab_bytes + VT_SLAB_SLACK + 127) & ~127) +
           (tile_is_deep(nfs, np) ? ((8 * nfs * lanes + 127) & ~127) : 0);
}

__global__ void __launch_bounds__(VS_PREP_THREADS) tile_prep2_kernel(BatchView b, ResultView r, Workspace ws, int32_t slab_cap, int32_t arena,
                                                                      TileHdr2 *hdr, FsTile *fs_tiles, int32_t *max_need, int32_t n_tiles,
                                                                      int32_t force_generic) {
    GCB_GRID_DEP();
    const int lane = lane_id();
    const int tile = (int)(blockIdx.x * (VS_PREP_THREADS / WARP) + (threadIdx.x >> 5));
    if (tile >= n_tiles || batch_is_malformed(ws.error_flag)) return;
    const TileDir t0 = ws.tile_dir[tile], t1 = ws.tile_dir[tile + 1];
    const int c0 = t0.c0, c1 = t1.c0;
    const int P0 = t0.p0, NP = t1.p0 - t0.p0;
    const int64_t slab_bytes = t1.slab0 - t0.slab0;
    TileHdr2 h;
    h.out_base0 = 0; h.slab0 = t0.slab0; h.slab_bytes = 0; h.p0 = P0; h.np = NP; h.nfs = 0; h.lanes = 1; h.common_l = 0;
    h.per_bundle = 32; h.n_bundles = 0;
    if (c0 >= c1 || NP == 0) {  // no cluster starts here / clusters without pairs emit nothing
        if (lane == 0) hdr[tile] = h;
        return;
    }
    if (NP > VS_MAX_PAIRS || slab_bytes > slab_cap || force_generic) {  // not a tile for the staged kernel
        if (lane == 0) {
            hdr[tile] = h;
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = (int32_t)tile;
            GCB_COUNT(1, 1);
        }
        return;
    }
    const int64_t out_base0 = ws.scan_block[c0 / SCAN_BLOCK] + ws.cluster_out_off[c0];
    FsTile *ft_out = fs_tiles + 2 * (int64_t)P0;
    uint32_t total = 0;
    bool nofit = false;
    int lmax = 1, common = 0;
    for (int base = 0; base < NP; base += WARP) {
        const int pos = base + lane;
        FsDesc fd[2];
        fd[0].mode = fd[1].mode = SIDE_NONE;
        fd[0].c = fd[1].c = c0;
        if (pos < NP) {  // slots that hold no family carry SIDE_NONE in side_mode and garbage in fs_desc
            const uint16_t modes = *(const uint16_t *)(ws.side_mode + 2 * (int64_t)(P0 + pos));
            if ((modes & 0xFF) != SIDE_NONE) fd[0] = ws.fs_desc[2 * (int64_t)(P0 + pos)];
            if ((modes >> 8) != SIDE_NONE) fd[1] = ws.fs_desc[2 * (int64_t)(P0 + pos) + 1];
        }
        const bool live0 = fd[0].mode != SIDE_NONE, live1 = fd[1].mode != SIDE_NONE;
        int64_t c_slab = 0, c_out = 0;
        if (live0 || live1) {
            const int c = live0 ? fd[0].c : fd[1].c;
            c_slab = ws.slab_off[c] - t0.slab0;
            c_out = ws.scan_block[c / SCAN_BLOCK] + ws.cluster_out_off[c] - out_base0;
        }
        const unsigned b0m = __ballot_sync(FULL, live0), b1m = __ballot_sync(FULL, live1), lt = (1u << lane) - 1u;
        int fidx = (int)total + __popc(b0m & lt) + __popc(b1m & lt);
        int lneed = 1;
        bool bad = false;
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
            ft.slot = P0 + pos;
            ft.reserved = 0;
            const int l = d.l_out;
            const int chunks = max((GCB_ALIGN4(l) + 15) >> 4, (GCB_ALIGN4((l + 1) >> 1) + 7) >> 3);
            if ((d.flags & FS_NOFIT) || (orel >> 2) > 0xFFFF || chunks > WARP) bad = true;
            if (out_base0 + orel + record_bytes(l) > r.out_capacity) {
                raise_error(ws.error_flag, GCB_ERR_CAPACITY);
                ft.mode = SIDE_NONE;  // keeps its place in the table but is never voted
            } else {
                // the absolute offset the caller reads (a tile that ends up with the generic kernel is listed as ~tile:
                // "offsets already absolute")
                r.groups[P0 + pos].out_off[side] = out_base0 + orel;
            }
            lneed = max(lneed, min(chunks, WARP));
            if (fidx < VS_MAX_FS) ft_out[fidx] = ft;
            if (fidx == 0) common = l;
            fidx++;
        }
        total += (uint32_t)(__popc(b0m) + __popc(b1m));
        nofit = nofit || __any_sync(FULL, bad);
        lmax = max(lmax, __reduce_max_sync(FULL, lneed));
        common = __reduce_max_sync(FULL, common);  // (only the lane that wrote entry 0 holds a non-zero value)
    }
    if (total > (uint32_t)VS_MAX_FS) nofit = true;
    const int32_t need = tile_smem_need((int32_t)min(total, (uint32_t)VS_MAX_FS), NP, (int32_t)slab_bytes, lmax);
    if (need > arena) nofit = true;  // (the host sized the window so that this is rare: a stage holds the tile or the generic kernel takes it)
    if (lane == 0) {
        if (nofit) {  // the generic kernel takes the tile
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = ~(int32_t)tile;
            GCB_COUNT(1, 1);
        } else if (total > 0) {
            h.out_base0 = out_base0;
            h.slab_bytes = (int32_t)slab_bytes;
            h.nfs = (int32_t)total;
            h.lanes = lmax;
            h.common_l = common;
            h.per_bundle = 32 / lmax;
            h.n_bundles = ((int32_t)total + h.per_bundle - 1) / h.per_bundle;
            atomicMax(max_need, need);
            GCB_COUNT(0, 1);
        }
        hdr[tile] = h;
    }
}

// ------------------------------------------------------------------------------------------------
// group.cpp:538-566 once every column is decided: a family side that collected more than five new mismatches keeps the
// template's bases and (rewritten) qualities.  The candidates were listed while the slow columns were decided (normally none).
constexpr int VQ_FINAL_THREADS = 128;
constexpr int VQ_FINAL_CTAS = 64;  // vote_rollback_kernel strides over its (normally empty) list

GCB_DEV void rollback_family_side(const BatchView &b, const ResultView &r, const Workspace &ws, const gcb_options &o, int64_t i) {
    const int slot = (int)(i >> 1), side = (int)(i & 1);
    const gcb_group_result *gr = r.groups + slot;
    if (gr->mismatch_inc[side] <= 5) return;
    const FsDesc d = ws.fs_desc[i];
    const VoteRead tv = ws.vote_reads[2 * (int64_t)d.mb + (int64_t)side * d.m + d.tmpl_k];
    const uint8_t *cb = b.payload + ws.slab_off[d.c];
    uint8_t *out = r.out_payload + gr->out_off[side];
    const int l_out = d.l_out, qbytes = GCB_ALIGN4(l_out);
    const uint8_t *tseq = cb + 4 * (int)tv.own_off4 + qbytes;
    for (int col = 0; col < l_out; col++) {
        int base, qual = 0, sc;
        fetch_ent(cb, tv, col, side, o, base, qual, sc);
        out[col] = (uint8_t)qual;
    }
    for (int k = 0; k < (l_out + 1) >> 1; k++) out[qbytes + k] = tseq[k];
}

__global__ void __launch_bounds__(VQ_FINAL_THREADS) vote_rollback_kernel(BatchView b, ResultView r, Workspace ws, gcb_options o, RollbackList rb,
                                                                         int32_t p0, int32_t p1) {
    GCB_GRID_DEP();
    if (batch_is_malformed(ws.error_flag)) return;
    const int n = *rb.count;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    if (n <= rb.cap) {
        for (int64_t k = tid; k < n; k += nthreads) rollback_family_side(b, r, ws, o, rb.list[k]);
    } else {  // the list overflowed: look at every family side of the view
        for (int64_t i = 2 * (int64_t)p0 + tid; i < 2 * (int64_t)p1; i += nthreads)
            if (ws.side_mode[i] != SIDE_NONE && ws.side_mode[i] != SIDE_COPY) rollback_family_side(b, r, ws, o, i);
    }
}

}  // namespace gcb
