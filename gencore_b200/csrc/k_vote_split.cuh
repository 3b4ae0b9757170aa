// k_vote_split.cuh — the vote split by column kind: Pair::computeScore (pair.cpp:88-172) fused with
// Group::makeConsensus (group.cpp:320-579), same arithmetic as k_vote_tiled.cuh, organised so that no CTA ever waits
// for a handful of threads that decide its few slow columns.
//
//   tile_prep2_kernel     (k_vote_staged.cuh) tile headers + compact family-side lists, once per batch.
//   vote_fast_kernel      one CTA per tile: three bulk asynchronous copies (cp.async.bulk -> UBLKCP, one mbarrier), ONE
//                         barrier, then every warp takes bundles of family sides and leaves.  A lane owns sixteen
//                         columns of one family side.  FAST columns (every voter shows the template's base, no mate
//                         disagrees, best quality >= moderateQuality: exactly group.cpp:421-427) are finished in the
//                         word.  For every other column the lane EMITS a record into a global queue: per read its
//                         quality, base nibble, its mate's quality and base nibble and where pair.cpp:121-170 puts the
//                         column (no overlap information / outside the overlap / mate base present / mate index out of
//                         range) — 4 bytes per read, 16 bytes of header.  One 64-bit atomic per bundle reserves the
//                         records of all its lanes (record count and words in one counter), thirty-two queues.
//   slow_columns_kernel   one thread per queued record, full warps: score per read (pair.cpp), three-bin register
//                         histogram (group.cpp:376-393), top-2 selection (group.cpp:395-417), the rules and the
//                         reference arbitration (group.cpp:419-525), patches the consensus record and adds to the
//                         family side's diff / mismatchInc (atomics on the result row; the fast kernel zeroed them).
//   vote_rollback_kernel  the > 5-mismatch rollback (group.cpp:538-566) of the family sides slow_columns_kernel listed.
//
// The uniform-family loop of the fast kernel keeps everything in the raw byte order of the payload: disagreement with
// the template and disagreement among the mates are OR-accumulated as XOR residues of the raw words (a mate differs
// from the template iff it differs from the template's own mate or that mate differs from the template), and the
// byte swap / funnel shift that aligns the mates with this lane's columns is applied once per bundle, after the
// loop.  Two reads per iteration feed a three-input 16-bit-lane maximum (VIMNMX3.U16x2); reads that do not vote are
// replaced by the template's own record (maximum and OR are idempotent), so the loop has no branches.
#pragma once

#include "k_vote_staged.cuh"

namespace gcb {

constexpr int VQ_NQ = 32;                      // slow-column queues; tile t uses queue t % VQ_NQ
constexpr uint32_t VQ_INVALID = 0xFFFFFFFFu;   // index entry of a reservation that did not fit
constexpr int VQ_SLOW_THREADS = 128;
constexpr int VQ_SLOW_PARTS = 96;              // CTAs per queue in slow_columns_kernel
constexpr int VQ_FINAL_THREADS = 128;
constexpr int VQ_FINAL_CTAS = 64;             // vote_rollback_kernel strides over its (normally empty) list

struct SlowQueues {
    unsigned long long *count;   // [VQ_NQ] records << 32 | words reserved so far (may run past the capacity)
    uint32_t *words;             // [VQ_NQ][cap_words] records (see SR_HDR_WORDS)
    uint32_t *index;             // [VQ_NQ][cap_recs] word offset of every record inside its queue, VQ_INVALID = none
    uint32_t cap_words, cap_recs;
    int32_t *rb_list;            // [rb_cap] family sides (2 * slot + side) whose mismatchInc went from 5 to 6: rollback candidates
    int32_t *rb_count;           // [1] entries appended to rb_list (may run past rb_cap: then every family side is checked)
    int32_t rb_cap;
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

#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ uint32_t lds32r(uint32_t addr) {  // ld.shared.u32 from a shared-window address
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
#else
inline uint32_t lds32r(uint32_t addr) { return *(const uint32_t *)(::simt::dyn_smem() + addr); }
#endif
#ifndef GCB_SIMT_CHECK
template <int IMM>
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
#else
template <int IMM>
inline uint32_t lds16(uint32_t addr) { return *(const uint16_t *)(::simt::dyn_smem() + addr + IMM); }
#endif

// byte flags (bit 7 of every byte) of columns 0-3 (a) and 4-7 (b) widened to a big-endian nibble mask
GCB_DEV uint32_t nibs_of_flags(uint32_t a, uint32_t b) {
    return (prmt(a, b, 0x8ACEu) & 0xF0F0F0F0u) | (prmt(a, b, 0x9BDFu) & 0x0F0F0F0Fu);
}
// bit 7 of every byte of x that is >= the byte of t4 (t4 = four copies of a threshold <= 128)
GCB_DEV uint32_t bytes_ge_flags(uint32_t x, uint32_t t4) { return (((x | 0x80808080u) - t4) | x) & 0x80808080u; }

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

// Pair::qual2score (pair.cpp:77-86) with the thresholds and the four scores in registers
struct ScoreTab {
    int hq, mq, lq, sh, sm, sl, sb;
    GCB_DEV explicit ScoreTab(const gcb_options &o)
        : hq(o.high_quality), mq(o.moderate_quality), lq(o.low_quality), sh(sc8(o.score_high)), sm(sc8(o.score_moderate)), sl(sc8(o.score_low)),
          sb(sc8(o.score_bad)) {}
    GCB_DEV int q2s(int q) const { return q >= hq ? sh : q >= mq ? sm : q >= lq ? sl : sb; }
};

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
    GCB_DEV void add(int base, int qual, int score) {
        total += score;
        // the bin that holds the code, else the first free one
        const bool h0 = b0 == base, h1 = b1 == base, h2 = b2 == base;
        const bool hit = h0 || h1 || h2;
        const bool u0 = h0 || (!hit && b0 < 0);
        const bool u1 = h1 || (!hit && b0 >= 0 && b1 < 0);
        const bool u2 = h2 || (!hit && b0 >= 0 && b1 >= 0 && b2 < 0);
        overflow = overflow || !(u0 || u1 || u2);
        if (u0) { b0 = base; c0++; s0 += score; q0 += qual; x0 = max(x0, qual); }
        if (u1) { b1 = base; c1++; s1 += score; q1 += qual; x1 = max(x1, qual); }
        if (u2) { b2 = base; c2++; s2 += score; q2 += qual; x2 = max(x2, qual); }
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

// base, rewritten quality and score of a queue entry: the same function of the same bytes as fetch_vote
GCB_DEV bool slow_decode(const ScoreTab &t, uint32_t ent, int side, int &base, int &qual, int &score) {
    if (!(ent & SE_VOTES)) return false;
    const int ql = (int)(ent & 0xFFu), mql = (int)((ent >> 8) & 0xFFu);
    base = (int)((ent >> 16) & 0xFu);
    const int mbase = (int)((ent >> 20) & 0xFu);
    const uint32_t st = (ent >> 24) & 3u;
    qual = ql;
    if (st == SE_MATE) {
        if (base == mbase) {  // pair.cpp:147-152
            score = sc8(t.q2s((ql + mql) / 2) + 4);
        } else {  // pair.cpp:153-169
            const int lq = side == 0 ? ql : mql, rq = side == 0 ? mql : ql;
            const bool mine = side == 0 ? lq >= rq : !(lq >= rq);
            score = mine ? sc8(t.q2s(lq >= rq ? lq - rq : rq - lq) - 3) : 0;
            qual = max(0, ql - mql);
        }
    } else {
        score = st == SE_PLAIN ? t.q2s(ql) : t.sm;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VS_MAX_THREADS, 3) vote_fast_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o,
                                                                      int32_t implied, const TileHdr2 *hdr, const FsTile *fs_tiles, SlowQueues sq) {
    GCB_DYN_SMEM(smem);
    uint64_t *bar = (uint64_t *)(smem + VS_OFF_BAR);
    int *s_next = (int *)(smem + VS_OFF_NEXT);
    FsTile *s_ft = (FsTile *)(smem + VS_OFF_FT);
    VoteRead *s_vr = (VoteRead *)(smem + VS_OFF_VR);
    uint8_t *slab = smem + VS_OFF_SLAB;
#define GCB_LDS32(off) (*(const uint32_t *)(smem + (off)))

    const int tid = (int)threadIdx.x, lane = lane_id();
    const TileHdr2 h = hdr[blockIdx.x];
    const int nfs = h.nfs;
    if (nfs == 0) return;
    if (tid == 0) {
        tile_barrier_init(bar);
        *s_next = 0;
        *(int *)(smem + VS_OFF_NSLOW) = 0;
        const uint32_t vb = 32u * (uint32_t)h.np, fb = 32u * (uint32_t)nfs;
        tile_expect(bar, (uint32_t)h.slab_bytes + vb + fb);
        if (h.slab_bytes > 0) tile_copy(slab, b.payload + h.slab0, (uint32_t)h.slab_bytes, bar);
        tile_copy(s_vr, ws.vote_reads + 2 * (int64_t)h.p0, vb, bar);
        tile_copy(s_ft, fs_tiles + 2 * (int64_t)h.p0, fb, bar);
    }
    __syncthreads();
    tile_wait(bar, 0);

    uint8_t *out0 = r.out_payload + h.out_base0;
    int *s_handed = (int *)(smem + VS_OFF_NSLOW);  // the tile went to the generic kernel (queue overflow)
    const int qi = (int)(blockIdx.x % VQ_NQ);
    uint32_t *q_words = sq.words + (size_t)qi * sq.cap_words;
    uint32_t *q_index = sq.index + (size_t)qi * sq.cap_recs;

    const uint32_t mod4 = 0x01010101u * (uint32_t)(o.moderate_quality & 0xFF);
    const uint32_t sbase = smem_base(smem);
    // (divisions of small numbers by multiply-and-shift: exact for numerators below 2^16 / divisor)
    const int L = h.lanes;                              // lanes per family side, 1..32
    const int S = (int)((32u * ((65535u / (unsigned)L) + 1u)) >> 16);  // family sides per bundle = 32 / L
    const int nb = (int)(((unsigned)(nfs + S - 1) * ((65535u / (unsigned)S) + 1u)) >> 16);
    const int sub = (int)(((unsigned)lane * ((65535u / (unsigned)L) + 1u)) >> 16), j = lane - sub * L;
    const int col0 = VT_CHUNK * j;
    const int common_l = h.common_l;  // the masks of the tile's usual record length are computed once
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
        const int cb = VS_OFF_SLAB + 4 * (int)ft.cbase4;  // byte offsets into the CTA's shared memory
        const int ento = VS_OFF_VR + 16 * (int)ft.ent0;
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
        if (mine && j == 0) {  // slow_columns_kernel adds to these
            gcb_group_result *gr = r.groups + ft.slot;
            const int sd = (ft.flags & FS_SIDE1) ? 1 : 0;
            gr->diff[sd] = 0;
            gr->mismatch_inc[sd] = 0;
        }
        // per-column maxima live in 16-bit lanes (VIMNMX.U16x2 / VIMNMX3.U16x2 are native, a per-byte maximum is seven
        // instructions): mo[k] tracks bytes 1 and 3 of quality word k in the high byte of each half, me[k] bytes 0 and 2
        uint32_t mo[4] = {0u, 0u, 0u, 0u}, me[4] = {0u, 0u, 0u, 0u}, dis0 = 0u, dis1 = 0u;
        if (ft.flags & FS_UNIFORM) {
            // hoisted geometry: every voter is read at the template's columns and meets its mate at the same offset,
            // and every mate's record lies at the same distance from its read's record (FS_UNIFORM)
            const int x = (int)tv.ov_own - col0;
            const int y = x - (int)tv.ov_mate;
            const int oa = max(max(0, x), y), oz = min(min(cm.nvote, x + (int)tv.ov_len), y + (int)tv.mate_l);
            const bool has_ov = tv.ov_len > 0 && oz > oa;
            const uint32_t om0 = has_ov ? nib_range(oa, oz) : 0u, om1 = has_ov ? nib_range(oa - 8, oz - 8) : 0u;
            const int ms = 0 - y, mw0 = ms >> 3;  // the lane's first mate column: word mw0 of the mate's bases, nibble ms & 7
            const unsigned msh = (unsigned)(ms & 7) * 4u;
            const uint32_t qbase = sbase + (uint32_t)(cb + col0), sdelta = (uint32_t)(qbytes - col0 + 8 * j);
            // mate words relative to the read's own quality chunk.  Words that hold no overlapped column read whatever
            // lies there (the CTA's shared memory: the lane's window starts at most two words before the mate's bases
            // and ends inside the slab's slack) and are masked by om0 / om1; lanes without overlap re-read their own chunk.
            const uint32_t mdelta = has_ov ? (uint32_t)(4 * ((int)tv.mate_off4 - (int)tv.own_off4) + GCB_ALIGN4(tv.mate_l) + 4 * mw0 - col0) : 0u;
            const uint32_t xt = tv.own_off4;
            const uint32_t qt = qbase + (xt << 2);
            const uint32_t t0 = lds32<0>(qt + sdelta), t1 = lds32<4>(qt + sdelta);               // the template's bases, raw
            const uint32_t a0 = lds32<0>(qt + mdelta), c0 = lds32<4>(qt + mdelta), e0 = lds32<8>(qt + mdelta);  // its mate's
            uint32_t d0 = 0u, d1 = 0u, da = 0u, dc = 0u, de = 0u;
            uint32_t ea = sbase + (uint32_t)ento;
            for (int e = 0; e < mmax; e += 2, ea += 32) {
                uint32_t xa = lds16<0>(ea), xb = lds16<16>(ea);
                xa = (e < m && xa != VR_NO_VOTE) ? xa : xt;
                xb = (e + 1 < m && xb != VR_NO_VOTE) ? xb : xt;
                const uint32_t qa = qbase + (xa << 2), qb = qbase + (xb << 2);
                const uint32_t qa0 = lds32<0>(qa), qa1 = lds32<4>(qa), qa2 = lds32<8>(qa), qa3 = lds32<12>(qa);
                const uint32_t qb0 = lds32<0>(qb), qb1 = lds32<4>(qb), qb2 = lds32<8>(qb), qb3 = lds32<12>(qb);
                const uint32_t ra0 = lds32<0>(qa + sdelta), ra1 = lds32<4>(qa + sdelta);
                const uint32_t rb0 = lds32<0>(qb + sdelta), rb1 = lds32<4>(qb + sdelta);
                const uint32_t ma = qa + mdelta, mb = qb + mdelta;
                const uint32_t aa = lds32<0>(ma), ca = lds32<4>(ma), ee = lds32<8>(ma);
                const uint32_t ab = lds32<0>(mb), cbb = lds32<4>(mb), eb = lds32<8>(mb);
                mo[0] = __vimax3_u16x2(mo[0], qa0, qb0); me[0] = __vimax3_u16x2(me[0], qa0 << 8, qb0 << 8);
                mo[1] = __vimax3_u16x2(mo[1], qa1, qb1); me[1] = __vimax3_u16x2(me[1], qa1 << 8, qb1 << 8);
                mo[2] = __vimax3_u16x2(mo[2], qa2, qb2); me[2] = __vimax3_u16x2(me[2], qa2 << 8, qb2 << 8);
                mo[3] = __vimax3_u16x2(mo[3], qa3, qb3); me[3] = __vimax3_u16x2(me[3], qa3 << 8, qb3 << 8);
                d0 |= (ra0 ^ t0) | (rb0 ^ t0);
                d1 |= (ra1 ^ t1) | (rb1 ^ t1);
                da |= (aa ^ a0) | (ab ^ a0);
                dc |= (ca ^ c0) | (cbb ^ c0);
                de |= (ee ^ e0) | (eb ^ e0);
            }
            dis0 = bswap32(d0);
            dis1 = bswap32(d1);
            if (has_ov) {  // pair.cpp:133-170: a base that differs from its mate's is never a fast column
                const uint32_t A = bswap32(da), C = bswap32(dc), E = bswap32(de);
                const uint32_t ta = bswap32(a0), tc = bswap32(c0), te = bswap32(e0);
                const uint32_t tb0 = bswap32(t0), tb1 = bswap32(t1);
                dis0 |= (__funnelshift_l(C, A, msh) | (tb0 ^ __funnelshift_l(tc, ta, msh))) & om0;
                dis1 |= (__funnelshift_l(E, C, msh) | (tb1 ^ __funnelshift_l(te, tc, msh))) & om1;
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
                if (v.ov_len > 0) {  // (subtractions only: see the ptxas note in k_vote_tiled.cuh)
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
        // ---- what the record gets: qualities = the maxima (fast columns), bases = the template's
        uint32_t slow0 = 0u, slow1 = 0u;
        if (mine) {
            uint32_t oq[4];
            if (ft.mode == SIDE_COPY) {  // group.cpp:73-77: the record itself
#pragma unroll
                for (int k = 0; k < 4; k++) oq[k] = col0 + 4 * k < qbytes ? GCB_LDS32(trec + col0 + 4 * k) : 0u;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) oq[k] = prmt(mo[k], me[k], 0x3715u) & cm.vb[k];  // (the hoisted loop read whole words)
                GCB_COUNT(2, cm.nvote);
                if (implied && len == l_out) {
                    const uint32_t lowq0 = nibs_of_flags(bytes_ge_flags(oq[0], mod4) ^ 0x80808080u, bytes_ge_flags(oq[1], mod4) ^ 0x80808080u);
                    const uint32_t lowq1 = nibs_of_flags(bytes_ge_flags(oq[2], mod4) ^ 0x80808080u, bytes_ge_flags(oq[3], mod4) ^ 0x80808080u);
                    slow0 = (dis0 | lowq0) & cm.vn0;
                    slow1 = (dis1 | lowq1) & cm.vn1;
                    // any differing bit marks the whole column
                    slow0 |= slow0 >> 1; slow0 |= slow0 >> 2; slow0 &= 0x11111111u;
                    slow1 |= slow1 >> 1; slow1 |= slow1 >> 2; slow1 &= 0x11111111u;
                } else {  // without `implied`, or with columns that are not voted, every column of the record is slow
                    slow0 = nibs_of_bytes(cm.rb[0], cm.rb[1]) & 0x11111111u;
                    slow1 = nibs_of_bytes(cm.rb[2], cm.rb[3]) & 0x11111111u;
                }
            }
            uint8_t *out = out0 + 4 * (int64_t)ft.out4;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (col0 + 4 * k < qbytes) *(uint32_t *)(out + col0 + 4 * k) = oq[k] & cm.rb[k];
            if (8 * j < sbytes) *(uint32_t *)(out + qbytes + 8 * j) = bswap32(tbe0 & cm.kn0);
            if (8 * j + 4 < sbytes) *(uint32_t *)(out + qbytes + 8 * j + 4) = bswap32(tbe1 & cm.kn1);
        }
        // ---- slow columns (one bit per column in slow0 / slow1): one reservation per bundle, every lane emits its own
        const int nslow = __popc(slow0) + __popc(slow1);
        if (__any_sync(FULL, nslow > 0)) {
            const uint32_t rec_words = slow_rec_words(ft.m);
            const uint32_t my_words = (uint32_t)nslow * rec_words;
            GCB_COUNT(3, nslow);
            // exclusive scans of the record counts and of the words over the warp (packed: records << 32 | words)
            unsigned long long mine64 = ((unsigned long long)(uint32_t)nslow << 32) | my_words, incl = mine64;
            for (int off = 1; off < WARP; off <<= 1) {
                const unsigned long long v = __shfl_up_sync(FULL, incl, off);
                if (lane >= off) incl += v;
            }
            const unsigned long long total = __shfl_sync(FULL, incl, WARP - 1);
            unsigned long long base64 = 0ull;
            if (lane == 0) base64 = atomicAdd(sq.count + qi, total);
            base64 = __shfl_sync(FULL, base64, 0);
            const uint32_t rec0 = (uint32_t)(base64 >> 32), word0 = (uint32_t)base64, T = (uint32_t)(total >> 32);
            const bool fits = (unsigned long long)rec0 + T <= sq.cap_recs && (unsigned long long)word0 + (uint32_t)total <= sq.cap_words;
            uint32_t ri = rec0 + (uint32_t)((incl - mine64) >> 32), wi = word0 + (uint32_t)(incl - mine64);
            if (!fits) {
                // the queue is full: the generic kernel redoes the whole tile from the payload (it runs after
                // slow_columns_kernel and vote_rollback_kernel); the reserved index entries are marked unused
                for (uint32_t i = (uint32_t)lane; i < T; i += WARP)
                    if (rec0 + i < sq.cap_recs) q_index[rec0 + i] = VQ_INVALID;
                if (lane == 0 && atomicExch(s_handed, 1) == 0) {
                    ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = ~(int32_t)blockIdx.x;
                    GCB_COUNT(1, 1);
                }
            } else if (nslow > 0) {
                const uint8_t *cbp = smem + cb;
                const VoteRead *ents = s_vr + ft.ent0;
                for (int wsel = 0; wsel < 2; wsel++) {
                    uint32_t sm = wsel ? slow1 : slow0;
                    while (sm != 0u) {
                        const int k = __clz((int)sm) >> 2;
                        sm &= ~(0xF0000000u >> (4 * k));
                        const int col = col0 + 8 * wsel + k;
                        uint32_t *rec = q_words + wi;
                        q_index[ri] = wi;
                        slow_write_header(rec, ft, col, h.out_base0 + 4 * (int64_t)ft.out4);
                        for (int e = 0; e < (int)ft.m; e++) rec[SR_HDR_WORDS + e] = slow_entry(cbp, ents[e], col);
                        ri++;
                        wi += rec_words;
                    }
                }
            }
        }
    }
#undef GCB_LDS32
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
                         const SlowQueues &sq, const uint32_t *rec) {
    const uint4 ha = ((const uint4 *)rec)[0], hc = ((const uint4 *)rec)[1];
    const uint32_t fsid = ha.x, w1 = ha.y, w2 = ha.z;
    const int col = (int)(w1 & 0xFFFFu), n = (int)(w1 >> 16), tmpl_k = (int)(w2 & 0xFFFFu);
    const uint32_t flags = w2 >> 16;
    const uint32_t *ents = rec + SR_HDR_WORDS;
    const int side = (int)(fsid & 1u);
    const int qbytes = GCB_ALIGN4((int)ha.w);
    uint8_t *out = r.out_payload + (int64_t)(((uint64_t)hc.y << 32) | hc.x);
    const int64_t ref_nib0 = (int64_t)(((uint64_t)hc.w << 32) | hc.z);
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
            gcb_group_result *gr = r.groups + (fsid >> 1);  // (fsid = 2 * slot + side; the fast kernel zeroed both counters)
            atomicAdd(&gr->diff[side], 1);
            if (d_mm != 0) {
                const int before = atomicAdd(&gr->mismatch_inc[side], d_mm);
                if (d_mm > 0 && before == 5) {  // more than five new mismatches so far: vote_rollback_kernel looks at the final count
                    const int k = atomicAdd(sq.rb_count, 1);
                    if (k < sq.rb_cap) sq.rb_list[k] = (int32_t)fsid;
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

__global__ void __launch_bounds__(VQ_SLOW_THREADS) slow_columns_kernel(BatchView b, ResultView r, GenomeView gv, gcb_options o, SlowQueues sq) {
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
#pragma unroll
        for (int step = VQ_NQ / 2; step > 0; step >>= 1)
            if (s_first[q + step] <= i) q += step;
        const uint32_t off = sq.index[(size_t)q * sq.cap_recs + (i - s_first[q])];
        if (off == VQ_INVALID) continue;
        slow_record(b, r, gv, o, tab, sq, sq.words + (size_t)q * sq.cap_words + off);
    }
}

// ------------------------------------------------------------------------------------------------
// group.cpp:538-566 once every column is decided: a family side that collected more than five new mismatches keeps the
// template's bases and (rewritten) qualities.  The candidates were listed by slow_columns_kernel (normally none).
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

__global__ void __launch_bounds__(VQ_FINAL_THREADS) vote_rollback_kernel(BatchView b, ResultView r, Workspace ws, gcb_options o, SlowQueues sq,
                                                                         int32_t p0, int32_t p1) {
    const int n = *sq.rb_count;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    if (n <= sq.rb_cap) {
        for (int64_t k = tid; k < n; k += nthreads) rollback_family_side(b, r, ws, o, sq.rb_list[k]);
    } else {  // the list overflowed: look at every family side of the view
        for (int64_t i = 2 * (int64_t)p0 + tid; i < 2 * (int64_t)p1; i += nthreads)
            if (ws.side_mode[i] != SIDE_NONE && ws.side_mode[i] != SIDE_COPY) rollback_family_side(b, r, ws, o, i);
    }
}

}  // namespace gcb
