// k_vote_tiled.cuh — the hot kernel: Pair::computeScore (pair.cpp:88-172) fused with
// Group::makeConsensus (group.cpp:320-579), eight template columns per thread.
//
// Work decomposition
//   CTA    = one tile: the consecutive clusters whose slab starts inside a window of the payload.
//            Three bulk asynchronous copies (cp.async.bulk -> UBLKCP, one mbarrier) stage the tile's
//            payload slab, its VoteRead table and its FsDesc table into shared memory; nothing else
//            of the batch is read on the common path.
//   thread = one item = eight consecutive columns of one (family, side): two 32-bit words of
//            qualities and one 32-bit word of 4-bit bases per read, compared and reduced with
//            SIMD-in-word arithmetic (no per-base loop, no warp shuffles).
// A column is FAST when every voting read shows the template's base there, no read disagrees with its
// mate inside the pair overlap, and the best quality reaches moderateQuality.  With the options for
// which that implies topScore >= baseScoreReq (`implied`, true for the reference's defaults) such a
// column is exactly group.cpp:421-427: second bin empty, only the quality (the maximum) is written.
// Every other column is SLOW: it is queued and decided by slow_column(), a literal per-read,
// per-bin restatement of group.cpp:376-525 that shares column_top()/column_arbitrate() with the
// generic kernel.  Tiles that do not fit the tables (huge clusters, thousands of tiny reads) are
// handed to score_vote_kernel (k_score_vote.cuh) through ws.generic_tiles.
#pragma once

#include "k_score_vote.cuh"

namespace gcb {

constexpr int VT_THREADS = 256;
constexpr int VT_MAX_PAIRS = VT_THREADS;  // pair positions of a tile: one thread each in the prologue
constexpr int VT_SLOW_CAP = 1024;         // queued slow columns; more are decided inline by their owner

// shared-memory map (bytes)
constexpr int VT_OFF_BAR = 0;
constexpr int VT_OFF_NSLOW = 8;
constexpr int VT_OFF_NOFIT = 12;
constexpr int VT_OFF_WSUM = 16;                                   // 8 warp totals of the chunk scan
constexpr int VT_OFF_CPO = 64;                                    // int32[VT_MAX_PAIRS + 2]
constexpr int VT_OFF_CHUNK0 = VT_OFF_CPO + 4 * (VT_MAX_PAIRS + 2 + 14);   // uint32[2*VT_MAX_PAIRS + 1]
constexpr int VT_OFF_ACC = VT_OFF_CHUNK0 + 4 * (2 * VT_MAX_PAIRS + 4);    // int32[2*VT_MAX_PAIRS]
constexpr int VT_OFF_SLOW = VT_OFF_ACC + 4 * 2 * VT_MAX_PAIRS;            // uint32[VT_SLOW_CAP]
constexpr int VT_OFF_FS = VT_OFF_SLOW + 4 * VT_SLOW_CAP;                  // FsDesc/FsTile[2*VT_MAX_PAIRS]
constexpr int VT_OFF_VR = VT_OFF_FS + 16 * 2 * VT_MAX_PAIRS;              // VoteRead[2*VT_MAX_PAIRS]
constexpr int VT_OFF_SLAB = (VT_OFF_VR + 16 * 2 * VT_MAX_PAIRS + 127) & ~127;
static_assert(VT_OFF_FS % 16 == 0 && VT_OFF_VR % 16 == 0, "bulk copy destinations are 16-byte aligned");
constexpr int VT_MAX_SLAB = 200 * 1024;  // cbase4 / out4 are 16-bit counts of 4-byte units

// FsDesc rewritten in place for the tile (still 16 bytes)
struct __align__(16) FsTile {
    uint16_t ent0;    // first VoteRead of this family side in the tile's table
    uint16_t m;
    uint16_t l_out;
    uint16_t len;
    uint16_t tmpl_k;
    uint8_t mode;
    uint8_t flags;
    uint16_t cbase4;  // the cluster's slab inside the staged tile, 4-byte units
    uint16_t out4;    // consensus record relative to the tile's first output byte, 4-byte units
};
static_assert(sizeof(FsTile) == 16 && sizeof(FsDesc) == 16 && sizeof(VoteRead) == 16, "table entries are 16 bytes");

// coverage counters of the CPU SIMT-check build (tests only): tiled tiles, generic tiles, fast columns, slow columns
#ifdef GCB_SIMT_CHECK
inline int64_t g_simt_counters[4] = {0, 0, 0, 0};
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
GCB_DEV uint32_t bswap32(uint32_t w) { return __byte_perm(w, 0, 0x0123); }
// columns [a, b) of a chunk, 0 <= a <= b <= 8, as a mask over the big-endian nibble word (column k = bits 28-4k..31-4k)
GCB_DEV uint32_t nib_range(int a, int b) {
    const uint32_t hi = a >= 8 ? 0u : (0xFFFFFFFFu >> (4 * a));
    const uint32_t lo = b >= 8 ? 0u : (0xFFFFFFFFu >> (4 * b));
    return hi & ~lo;
}
// bytes [a, b) of a little-endian word, arguments clamped to 0..4
GCB_DEV uint32_t byte_range(int a, int b) {
    a = max(a, 0);
    b = min(b, 4);
    if (b <= a) return 0u;
    const uint32_t lo = 0xFFFFFFFFu << (8 * a);  // a <= 3 here
    const uint32_t hi = b >= 4 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu << (8 * b));
    return lo & hi;
}
// eight qualities at read positions rp0..rp0+7 of a record whose quality area holds qbytes bytes; outside reads as 0
GCB_DEV void fetch8q(const uint8_t *rec, int qbytes, int rp0, uint32_t &q0, uint32_t &q1) {
    const uint32_t *p = (const uint32_t *)rec;
    const int nw = qbytes >> 2, w0 = rp0 >> 2;
    const unsigned sh = (unsigned)(rp0 & 3) * 8u;
    const uint32_t a = (unsigned)w0 < (unsigned)nw ? p[w0] : 0u;
    const uint32_t c = (unsigned)(w0 + 1) < (unsigned)nw ? p[w0 + 1] : 0u;
    const uint32_t d = (unsigned)(w0 + 2) < (unsigned)nw ? p[w0 + 2] : 0u;
    q0 = __funnelshift_r(a, c, sh);
    q1 = __funnelshift_r(c, d, sh);
}
// eight base codes at read positions rp0..rp0+7 as a big-endian nibble word; seq area holds sbytes bytes
GCB_DEV uint32_t fetch8b(const uint8_t *seq, int sbytes, int rp0) {
    const uint32_t *p = (const uint32_t *)seq;
    const int nw = sbytes >> 2, w0 = rp0 >> 3;
    const unsigned sh = (unsigned)(rp0 & 7) * 4u;
    const uint32_t a = (unsigned)w0 < (unsigned)nw ? bswap32(p[w0]) : 0u;
    const uint32_t c = (unsigned)(w0 + 1) < (unsigned)nw ? bswap32(p[w0 + 1]) : 0u;
    return __funnelshift_l(c, a, sh);
}

// base, rewritten quality and score of one read at template column i: pair.cpp:88-172 for one base,
// from the staged tables (the same function of the same bytes as fetch_base in k_score_vote.cuh)
GCB_DEV bool fetch_ent(const uint8_t *cb, const VoteRead &v, int i, int side, const gcb_options &o, int &base, int &qual, int &score) {
    if (v.own_l == 0) return false;
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
    const Workspace *ws;
    const GenomeView *gv;
    const gcb_options *o;
    const uint8_t *slab;
    const VoteRead *vr;
    const FsTile *ft;
    const int32_t *cpo;  // cluster_pair_off[c0 .. c1]
    int32_t *acc;
    uint8_t *out0;       // out_payload + the tile's first output byte
    int c0, nc, p0;
};

// cluster (relative to c0) that owns pair position `pos`
GCB_DEV int cluster_of(const int32_t *cpo, int nc, int pos) {
    int lo = 0, hi = nc;  // first index in (0, nc] whose offset is > pos; cpo[nc] > pos always
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cpo[mid] > pos) hi = mid;
        else lo = mid + 1;
    }
    return lo - 1;
}

// One column decided the long way: group.cpp:376-525.  Thread-local.
GCB_DEV void slow_column(const TileCtx &t, int f, int col) {
    const gcb_options &o = *t.o;
    const FsTile ft = t.ft[f];
    const int side = f & 1, slot = t.p0 + (f >> 1);
    const uint8_t *cb = t.slab + 4 * (int)ft.cbase4;
    const VoteRead *ents = t.vr + ft.ent0;
    const VoteRead tv = ents[ft.tmpl_k];
    const int l_out = ft.l_out, qbytes = GCB_ALIGN4(l_out);
    uint8_t *out = t.out0 + 4 * (int64_t)ft.out4;
    int obase = 0, oqual = 0, sc;
    fetch_ent(cb, tv, col, side, o, obase, oqual, sc);
    if (col >= ft.len) {  // beyond the voted columns the record keeps what it held (rewritten qualities)
        out[col] = (uint8_t)oqual;
        return;
    }
    SparseBins bins;
    bins.init();
    for (int e = 0; e < ft.m; e++) {
        int base, qual, score;
        if (fetch_ent(cb, ents[e], col, side, o, base, qual, score)) bins.add(base, qual, score);
    }
    VoteBin obs[16];
    int nobs = 0, total = bins.total;
    if (bins.overflow) {  // four or more distinct codes: full histogram
        int cnt[16], scs[16], qls[16], mxq[16];
        for (int k = 0; k < 16; k++) cnt[k] = scs[k] = qls[k] = mxq[k] = 0;
        total = 0;
        for (int e = 0; e < ft.m; e++) {
            int base, qual, score;
            if (!fetch_ent(cb, ents[e], col, side, o, base, qual, score)) continue;
            cnt[base]++;
            scs[base] += score;
            qls[base] += qual;
            mxq[base] = max(mxq[base], qual);
            total += score;
        }
        for (int k = 0; k < 16; k++)
            if (cnt[k] > 0) {
                obs[nobs].base = k; obs[nobs].cnt = cnt[k]; obs[nobs].score = scs[k]; obs[nobs].qual = qls[k]; obs[nobs].maxq = mxq[k];
                nobs++;
            }
    } else {
        for (int k = 0; k < 3; k++)
            if (bins.s[k].base >= 0) obs[nobs++] = bins.s[k];
    }
    const ColumnTop top = column_top(o, obs, nobs, total);
    int new_base = obase, new_qual;
    if (top.fast) {
        new_qual = top.top.maxq;  // group.cpp:422-426: the base is NOT written
    } else {
        int ref4 = 0;
        const int tmpl = t.r->groups[slot].tmpl_read[side];
        const gcb_read_desc od = t.b->reads[tmpl];
        if (od.isize != 0 && t.gv->packed4) {  // group.cpp:362-367 + reference.cpp:33-71, group.cpp:430-439
            const int c = t.c0 + cluster_of(t.cpo, t.nc, slot);
            const int contig = t.b->cluster_ref[c];
            if (contig >= 0 && contig < t.gv->n_contigs) {
                const uint32_t *ocig = t.b->cigar + od.cigar_off;
                const int64_t span = (int64_t)get_ref_offset(ocig, od.n_cigar, ft.len - 1) + 1;
                const int64_t clen = t.gv->contig_len[contig];
                if ((int64_t)od.pos + span < clen) {
                    const int refpos = get_ref_offset(ocig, od.n_cigar, col);
                    const int64_t gp = (int64_t)od.pos + refpos;
                    if (refpos >= 0 && gp < clen) {
                        const uint8_t two = t.gv->packed4[t.gv->contig_off[contig] + (gp >> 1)];
                        ref4 = genome_nibble_to_bam((gp & 1) ? (two >> 4) : (two & 0xF));
                    }
                }
            }
        }
        int rbq = 0;
        bool any_high = false;
        if (top.need_ref && ref4 != 0) {
            int rmax = 0;
            for (int k = 0; k < nobs; k++)
                if (obs[k].base == ref4) rmax = obs[k].maxq;
            if (rmax >= 128) {  // `char refBaseQual` wraps: the scan order matters (group.cpp:474-490): template first
                int tb, tq, ts;
                if (fetch_ent(cb, tv, col, side, o, tb, tq, ts) && tb == ref4) {
                    if (tq > rbq) rbq = sc8(tq);
                    if (tq >= o.high_quality) any_high = true;
                }
                for (int e = 0; e < ft.m; e++) {
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
        int d_diff = 0, d_mm = 0;
        if (obase != co.base) {  // group.cpp:509-524
            new_base = co.base;
            d_diff = 1;
            if (ref4 != 0) {
                if (obase == ref4) d_mm = 1;
                else if (co.base == ref4) d_mm = -1;
            }
            atomicAdd(t.acc + f, d_diff + d_mm * 65536);
            const int byte = col >> 1;
            const unsigned delta = ((unsigned)(obase ^ new_base) & 0xFu) << ((col & 1) ? 0 : 4);
            atomicXor((unsigned *)(out + qbytes + (byte & ~3)), delta << (8 * (byte & 3)));
        }
        new_qual = co.qual;
    }
    out[col] = (uint8_t)new_qual;
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
        fetch_ent(cb, tv, col, f & 1, *t.o, base, qual, sc);
        out[col] = (uint8_t)qual;
    }
    for (int k = 0; k < (l_out + 1) >> 1; k++) out[qbytes + k] = tseq[k];
}

__global__ void __launch_bounds__(VT_THREADS) vote_tiled_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o,
                                                                int32_t slab_cap, int32_t implied) {
    GCB_DYN_SMEM(smem);
    uint64_t *bar = (uint64_t *)(smem + VT_OFF_BAR);
    int *s_nslow = (int *)(smem + VT_OFF_NSLOW);
    int *s_nofit = (int *)(smem + VT_OFF_NOFIT);
    uint32_t *s_wsum = (uint32_t *)(smem + VT_OFF_WSUM);
    int32_t *s_cpo = (int32_t *)(smem + VT_OFF_CPO);
    uint32_t *s_chunk0 = (uint32_t *)(smem + VT_OFF_CHUNK0);
    int32_t *s_acc = (int32_t *)(smem + VT_OFF_ACC);
    uint32_t *s_slow = (uint32_t *)(smem + VT_OFF_SLOW);
    FsDesc *s_fd = (FsDesc *)(smem + VT_OFF_FS);
    FsTile *s_ft = (FsTile *)(smem + VT_OFF_FS);
    VoteRead *s_vr = (VoteRead *)(smem + VT_OFF_VR);
    uint8_t *slab = smem + VT_OFF_SLAB;

    const int tid = (int)threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const TileDir t0 = ws.tile_dir[blockIdx.x], t1 = ws.tile_dir[blockIdx.x + 1];
    const int c0 = t0.c0, c1 = t1.c0;
    if (c0 >= c1) return;
    const int P0 = t0.p0, NP = t1.p0 - t0.p0, NC = c1 - c0;
    const int64_t slab_bytes = t1.slab0 - t0.slab0;
    if (NP > VT_MAX_PAIRS || NC > VT_MAX_PAIRS || slab_bytes > slab_cap) {  // not a tile for this kernel
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
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t tb = 32u * (uint32_t)NP;
        tile_expect(bar, (uint32_t)slab_bytes + 2u * tb);
        if (slab_bytes > 0) tile_copy(slab, b.payload + t0.slab0, (uint32_t)slab_bytes, bar);
        tile_copy(s_vr, ws.vote_reads + 2 * (int64_t)P0, tb, bar);
        tile_copy(s_fd, ws.fs_desc + 2 * (int64_t)P0, tb, bar);
    }
    for (int i = tid; i <= NC; i += VT_THREADS) s_cpo[i] = b.cluster_pair_off[c0 + i];
    const int64_t out_base0 = ws.scan_block[c0 / SCAN_BLOCK] + ws.cluster_out_off[c0];
    __syncthreads();

    // ---- prologue: one thread per pair position = per possible family slot
    int ci = 0;
    int64_t c_slab = 0, c_out = 0;
    if (tid < NP) {
        ci = cluster_of(s_cpo, NC, P0 + tid);
        const int c = c0 + ci;
        c_slab = ws.slab_off[c] - t0.slab0;
        c_out = ws.scan_block[c / SCAN_BLOCK] + ws.cluster_out_off[c] - out_base0;
        s_acc[2 * tid] = 0;
        s_acc[2 * tid + 1] = 0;
    }
    tile_wait(bar, 0);
    uint32_t nch[2] = {0u, 0u};
    if (tid < NP) {
        for (int side = 0; side < 2; side++) {
            const FsDesc fd = s_fd[2 * tid + side];
            FsTile ft;
            ft.ent0 = (uint16_t)(2 * (s_cpo[ci] - P0 + (int)fd.mb_rel) + side * (int)fd.m);
            ft.m = fd.m;
            ft.l_out = fd.l_out;
            ft.len = fd.len;
            ft.tmpl_k = fd.tmpl_k;
            ft.mode = fd.mode;
            ft.flags = fd.flags;
            ft.cbase4 = (uint16_t)(c_slab >> 2);
            const int64_t orel = c_out + fd.out_rel;
            ft.out4 = (uint16_t)(orel >> 2);
            if (fd.mode != SIDE_NONE) {
                if ((fd.flags & FS_NOFIT) || (orel >> 2) > 0xFFFF) *s_nofit = 1;
                const int l = fd.l_out;
                if (out_base0 + orel + record_bytes(l) > r.out_capacity) {
                    raise_error(ws.error_flag, GCB_ERR_CAPACITY);
                    ft.mode = SIDE_NONE;
                } else {
                    nch[side] = (uint32_t)max((GCB_ALIGN4(l) + 7) >> 3, GCB_ALIGN4((l + 1) >> 1) >> 2);
                }
            }
            s_ft[2 * tid + side] = ft;
        }
    }
    // exclusive scan of the chunk counts over the family sides of the tile
    {
        const uint32_t mine = nch[0] + nch[1];
        uint32_t incl = mine;
        for (int off = 1; off < WARP; off <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == WARP - 1) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t pre = incl - mine;
        for (int w = 0; w < warp; w++) pre += s_wsum[w];
        if (tid < NP) {
            s_chunk0[2 * tid] = pre;
            s_chunk0[2 * tid + 1] = pre + nch[0];
        }
        if (tid == VT_THREADS - 1) s_chunk0[2 * NP] = pre + mine;
        __syncthreads();
    }
    if (*s_nofit) {  // a field overflowed its table slot: the generic kernel takes the tile
        if (tid == 0) {
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = (int32_t)blockIdx.x;
            GCB_COUNT(1, 1);
        }
        return;
    }
    if (tid == 0) GCB_COUNT(0, 1);

    TileCtx t;
    t.b = &b; t.r = &r; t.ws = &ws; t.gv = &gv; t.o = &o;
    t.slab = slab; t.vr = s_vr; t.ft = s_ft; t.cpo = s_cpo; t.acc = s_acc;
    t.out0 = r.out_payload + out_base0;
    t.c0 = c0; t.nc = NC; t.p0 = P0;

    const uint32_t mod4 = 0x01010101u * (uint32_t)(o.moderate_quality & 0xFF);
    const int nfs = 2 * NP;
    const int total = (int)s_chunk0[nfs];
    for (int item = tid; item < total; item += VT_THREADS) {
        int f = 0;
        {
            int hi = nfs;  // last f with chunk0[f] <= item
            while (hi - f > 1) {
                const int mid = (f + hi) >> 1;
                if ((int)s_chunk0[mid] <= item) f = mid;
                else hi = mid;
            }
        }
        const FsTile ft = s_ft[f];
        const int j = item - (int)s_chunk0[f];
        const int col0 = 8 * j;
        const int l_out = ft.l_out, len = ft.len;
        const int qbytes = GCB_ALIGN4(l_out), sbytes = GCB_ALIGN4((l_out + 1) >> 1);
        const uint8_t *cb = slab + 4 * (int)ft.cbase4;
        const VoteRead *ents = s_vr + ft.ent0;
        const uint8_t *trec = cb + 4 * (int)ents[ft.tmpl_k].own_off4;
        const uint32_t tbe = 4 * j < sbytes ? bswap32(*(const uint32_t *)(trec + qbytes + 4 * j)) : 0u;
        const int nv = min(max(l_out - col0, 0), 8);                           // columns of the record in this chunk
        const int nk = min(max(2 * ((l_out + 1) >> 1) - col0, 0), 8);         // nibbles the record keeps (odd tail included)
        const uint32_t qm0 = byte_range(0, nv), qm1 = byte_range(0, nv - 4);
        uint32_t oq0, oq1, inline_slow = 0u;
        if (ft.mode == SIDE_COPY) {  // group.cpp:73-77: the record itself
            oq0 = col0 < qbytes ? *(const uint32_t *)(trec + col0) : 0u;
            oq1 = col0 + 4 < qbytes ? *(const uint32_t *)(trec + col0 + 4) : 0u;
        } else {
            uint32_t mq0 = 0u, mq1 = 0u, dis = 0u;
            for (int e = 0; e < (int)ft.m; e++) {
                const VoteRead v = ents[e];
                if (v.own_l == 0) continue;
                const int rp0 = col0 + v.shift;
                const int a = max(0, -rp0), z = min(8, min((int)v.own_l - rp0, len - col0));
                if (z <= a) continue;
                const uint8_t *rec = cb + 4 * (int)v.own_off4;
                const int rq = GCB_ALIGN4(v.own_l);
                uint32_t q0, q1, be;
                if (v.shift == 0) {
                    q0 = *(const uint32_t *)(rec + col0);
                    q1 = col0 + 4 < rq ? *(const uint32_t *)(rec + col0 + 4) : 0u;
                    be = bswap32(*(const uint32_t *)(rec + rq + 4 * j));
                } else {
                    fetch8q(rec, rq, rp0, q0, q1);
                    be = fetch8b(rec + rq, GCB_ALIGN4((v.own_l + 1) >> 1), rp0);
                }
                if (a != 0 || z != 8) {
                    q0 &= byte_range(a, z);
                    q1 &= byte_range(a - 4, z - 4);
                }
                mq0 = __vmaxu4(mq0, q0);
                mq1 = __vmaxu4(mq1, q1);
                dis |= (be ^ tbe) & nib_range(a, z);
                if (v.ov_len > 0) {  // pair.cpp:133-170: a base that differs from its mate's is never a fast column
                    // chunk column k pairs own index rp0+k with mate index k - y.  (Written with subtractions only:
                    // ptxas 12.9 dropped the negation when it folded max(a, max(x, -t)) into one VIMNMX3 on sm_100a —
                    // the PTX was right, the SASS and the B200 were not; profiles/r01_notes.md has the listing.)
                    const int x = (int)v.ov_own - rp0;   // first chunk column inside the overlap window
                    const int y = x - (int)v.ov_mate;    // first chunk column whose mate index is >= 0
                    const int oa = max(max(a, x), y);
                    const int oz = min(min(z, x + (int)v.ov_len), y + (int)v.mate_l);
                    if (oz > oa) {
                        const uint8_t *mrec = cb + 4 * (int)v.mate_off4;
                        const uint32_t mbe = fetch8b(mrec + GCB_ALIGN4(v.mate_l), GCB_ALIGN4((v.mate_l + 1) >> 1), 0 - y);
                        dis |= (be ^ mbe) & nib_range(oa, oz);
                    }
                }
            }
            oq0 = mq0;
            oq1 = mq1;
            GCB_COUNT(2, min(max(len - col0, 0), 8));
            // which columns are not fast
            const int nvote = min(max(len - col0, 0), 8);
            const uint32_t vq0 = byte_range(0, nvote), vq1 = byte_range(0, nvote - 4);
            const bool all_fast = implied && len == l_out && dis == 0u && ((__vcmpgeu4(mq0, mod4) & vq0) == vq0) &&
                                  ((__vcmpgeu4(mq1, mod4) & vq1) == vq1);
            if (!all_fast) {
                const int ncheck = (implied && len == l_out) ? nvote : nv;  // without `implied` (or with unvoted columns) every column is slow
                for (int k = 0; k < ncheck; k++) {
                    bool slow = !(implied && len == l_out);
                    if (!slow) {
                        const uint32_t q = ((k < 4 ? mq0 : mq1) >> (8 * (k & 3))) & 0xFFu;
                        slow = ((dis >> (28 - 4 * k)) & 0xFu) != 0u || (int)q < o.moderate_quality;
                    }
                    if (!slow) continue;
                    GCB_COUNT(3, 1);
                    const int idx = atomicAdd(s_nslow, 1);
                    if (idx < VT_SLOW_CAP) s_slow[idx] = ((uint32_t)f << 16) | (uint32_t)(col0 + k);
                    else inline_slow |= 1u << k;  // queue full: this thread owns the chunk's words and decides the column itself
                }
            }
        }
        uint8_t *out = t.out0 + 4 * (int64_t)ft.out4;
        if (col0 < qbytes) *(uint32_t *)(out + col0) = oq0 & qm0;
        if (col0 + 4 < qbytes) *(uint32_t *)(out + col0 + 4) = oq1 & qm1;
        if (4 * j < sbytes) *(uint32_t *)(out + qbytes + 4 * j) = bswap32(tbe & nib_range(0, nk));
        for (int k = 0; inline_slow != 0u; k++, inline_slow >>= 1)
            if (inline_slow & 1u) slow_column(t, f, col0 + k);
    }
    __syncthreads();
    {
        const int n = min(*s_nslow, VT_SLOW_CAP);
        for (int i = tid; i < n; i += VT_THREADS) slow_column(t, (int)(s_slow[i] >> 16), (int)(s_slow[i] & 0xFFFFu));
    }
    __syncthreads();
    if (tid < NP) {  // per family side: diff, mismatchInc, rollback, absolute output offset
        for (int side = 0; side < 2; side++) {
            const int f = 2 * tid + side;
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
}

}  // namespace gcb
