// k_vote_staged.cuh — the vote with its per-tile bookkeeping done once per batch and its slow columns decided one
// thread per column: Pair::computeScore (pair.cpp:88-172) fused with Group::makeConsensus (group.cpp:320-579), the
// same arithmetic and the same lane code as k_vote_tiled.cuh.
//
//   tile_prep2_kernel    one CTA per tile, after the output-offset scan: compacts the tile's live family sides into
//                        FsTile entries in global memory and leaves a 48-byte tile header (payload slab, pair range,
//                        output base, lanes per family side); writes the absolute output offsets the caller reads;
//                        hands tiles that do not fit to the generic kernel.
//   vote_staged_kernel   one CTA per tile.  Every thread reads the header (one broadcast load); thread 0 issues
//                        three bulk asynchronous copies (cp.async.bulk -> UBLKCP, one mbarrier): payload slab,
//                        VoteRead table, FsTile list.  No per-tile prologue, no scan: after ONE barrier the warps
//                        take bundles of family sides (sixteen columns per lane, hoisted masks for uniform
//                        families, 16-bit-lane maxima); after a second barrier the queued slow columns are decided
//                        one thread per column (slow_column: for a uniform family the overlap geometry of the
//                        column is computed once, every read then costs two or four shared-memory bytes and a
//                        three-bin register histogram); after a third, one thread per family side writes
//                        diff / mismatchInc and performs the > 5-mismatch rollback.
#pragma once

#include "k_vote_pipe.cuh"

namespace gcb {

constexpr int VS_MAX_THREADS = 256;
constexpr int VS_MAX_PAIRS = 256;   // pair positions of a tile
constexpr int VS_MAX_FS = 192;      // family sides of a tile
constexpr int VS_SLOW_CAP = 384;    // queued slow columns; more are decided inline by their owner
constexpr int VS_PREP_THREADS = 128;  // tile_prep2_kernel: four tiles per CTA, one warp each

struct __align__(16) TileHdr2 {
    int64_t out_base0;   // first output byte of the tile
    int64_t slab0;       // payload offset of the tile's slab
    int32_t slab_bytes;
    int32_t p0, np;      // pair positions
    int32_t nfs;         // live family sides (FsTile entries at fs_tiles[2*p0 ..]); 0 = nothing for this kernel
    int32_t lanes;       // lanes per family side: the tile's widest record in 16-column chunks
    int32_t common_l;    // l_out of the first family side (mask set computed once per tile)
    int32_t per_bundle;  // family sides per warp pass = 32 / lanes
    int32_t n_bundles;   // ceil(nfs / per_bundle)
};
static_assert(sizeof(TileHdr2) == 48, "tile header size");

// shared-memory map (bytes)
constexpr int VS_OFF_BAR = 0;
constexpr int VS_OFF_NSLOW = 8;
constexpr int VS_OFF_NEXT = 12;
constexpr int VS_OFF_ACC = 64;                                   // int32[VS_MAX_FS]
constexpr int VS_OFF_SLOW = VS_OFF_ACC + 4 * VS_MAX_FS;          // uint32[VS_SLOW_CAP]
constexpr int VS_OFF_FT = VS_OFF_SLOW + 4 * VS_SLOW_CAP;         // FsTile[VS_MAX_FS]
constexpr int VS_OFF_VR = VS_OFF_FT + 32 * VS_MAX_FS;            // VoteRead[2*VS_MAX_PAIRS]
constexpr int VS_OFF_SLAB = (VS_OFF_VR + 32 * VS_MAX_PAIRS + 127) & ~127;
static_assert(VS_OFF_FT % 16 == 0 && VS_OFF_VR % 16 == 0, "16-byte aligned tables");

// ------------------------------------------------------------------------------------------------
// One WARP per tile, 32 pair positions per pass: what the tiled kernel's prologue computes, once per batch.  The live family
// sides are compacted with ballots (no shared memory, no CTA barrier); many tiles per SM keep their chains of dependent loads
// (directory -> side modes -> family-side descriptors -> cluster offsets) in flight at once.
__global__ void __launch_bounds__(VS_PREP_THREADS) tile_prep2_kernel(BatchView b, ResultView r, Workspace ws, int32_t slab_cap, TileHdr2 *hdr,
                                                                      FsTile *fs_tiles, int32_t *max_need, int32_t n_tiles) {
    const int lane = lane_id();
    const int tile = (int)(blockIdx.x * (VS_PREP_THREADS / WARP) + (threadIdx.x >> 5));
    if (tile >= n_tiles) return;
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
    if (NP > VS_MAX_PAIRS || slab_bytes > slab_cap) {  // not a tile for the staged kernels
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
            // shared memory the tile takes in the ring kernel: family-side list, VoteRead table, slab + slack, each rounded to 128
            atomicMax(max_need, (int32_t)(((32 * (int32_t)total + 127) & ~127) + ((32 * NP + 127) & ~127) +
                                          (((int32_t)slab_bytes + VT_SLAB_SLACK + 127) & ~127)));
            GCB_COUNT(0, 1);
        }
        hdr[tile] = h;
    }
}

// ------------------------------------------------------------------------------------------------
// Slow columns that the register path below does not take (family sides that are not uniform, columns beyond the
// voted length, a fourth distinct code in one column) and records rolled back: out of line, one copy per kernel.
__device__ __noinline__ void slow_column_general(const TileCtx &t, int f, int col) { slow_thread(t, f, col); }
__device__ __noinline__ void rollback_record_general(const TileCtx &t, int f) { rollback_record(t, f); }

// group.cpp:376-525 for one slow column by one thread.  For a uniform family side (every voter has the template's
// length, no column shift, the same overlap window) what pair.cpp:121-170 needs to know about the column — inside the
// overlap or not, the mate index — is computed once; each read then is its quality byte, its base nibble and, inside
// the overlap, its mate's, added to a three-bin register histogram.  The two scans of group.cpp:395-417 are a top-2
// selection over the three bins and the two largest codes nobody showed (bin_key order), in registers.
GCB_DEV void slow_column(const TileCtx &t, int f, int col) {
    const FsTile ft = t.ft[f];
    if (!(ft.flags & FS_UNIFORM) || col >= (int)ft.len) {
        slow_column_general(t, f, col);
        return;
    }
    const gcb_options &o = *t.o;
    const uint8_t *cb = t.slab + 4 * (int)ft.cbase4;
    const VoteRead *ents = t.vr + ft.ent0;
    const VoteRead tv = ents[ft.tmpl_k];
    const int side = fs_side(ft);
    const bool info = tv.ov_len != VR_NO_OVERLAP_INFO;
    const int k = col - (int)tv.ov_own, mp = (int)tv.ov_mate + k;
    const bool inwin = info && k >= 0 && k < (int)tv.ov_len;
    const bool mvalid = inwin && mp >= 0 && mp < (int)tv.mate_l;
    const bool plain = info && !inwin;  // pair.cpp:121-131: outside the overlap the score follows the quality
    const int moderate = sc8(o.score_moderate);
    const int soff = GCB_ALIGN4(ft.l_out) + (col >> 1), nsh = (col & 1) ? 0 : 4;
    const int mpi = mvalid ? mp : 0;
    const int msoff = GCB_ALIGN4(tv.mate_l) + (mpi >> 1), mnsh = (mpi & 1) ? 0 : 4;
    SparseBins bins;
    bins.init();
    for (int e = 0; e < (int)ft.m; e++) {
        const uint32_t w = *(const uint32_t *)(ents + e);  // own_off4 | mate_off4 << 16
        if ((w & 0xFFFFu) == VR_NO_VOTE) continue;
        const uint8_t *rec = cb + 4 * (int)(w & 0xFFFFu);
        int ql = rec[col];
        const int base = (rec[soff] >> nsh) & 0xF;
        int score;
        if (mvalid) {
            const uint8_t *mrec = cb + 4 * (int)(w >> 16);
            const int mql = mrec[mpi];
            const int mbase = (mrec[msoff] >> mnsh) & 0xF;
            if (base == mbase) {  // pair.cpp:147-152
                score = sc8(qual2score_sel(o, (ql + mql) / 2) + 4);
            } else {  // pair.cpp:153-169
                const int lq = side == 0 ? ql : mql, rq = side == 0 ? mql : ql;
                const bool mine = side == 0 ? lq >= rq : !(lq >= rq);
                score = mine ? sc8(qual2score_sel(o, lq >= rq ? lq - rq : rq - lq) - 3) : 0;
                ql = max(0, ql - mql);
            }
        } else {
            score = plain ? qual2score_sel(o, ql) : moderate;
        }
        bins.add(base, ql, score);
    }
    if (bins.overflow) {
        slow_column_general(t, f, col);
        return;
    }
    // top and second: every bin competes with its (score, quality sum, code) key; the codes nobody showed compete
    // with (0, 0, code), of which only the two largest can place
    unsigned freemask = 0xFFFFu;
    unsigned long long key[3];
    uint32_t acgt = 0;
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
        const int bb = bins.s[kk].base;
        const bool have = bb >= 0;
        key[kk] = have ? bin_key(bins.s[kk].score, bins.s[kk].qual, bb) : 0ull;
        if (have) freemask &= ~(1u << bb);
        if (have && (bb == 1 || bb == 2 || bb == 4 || bb == 8)) acgt |= (uint32_t)bins.s[kk].maxq << (bb == 1 ? 0 : bb == 2 ? 8 : bb == 4 ? 16 : 24);
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
    ColumnTop ct;
    ct.top = bins.s[0].base == tb ? bins.s[0] : bins.s[1].base == tb ? bins.s[1] : bins.s[2].base == tb ? bins.s[2] : none;
    ct.sec = bins.s[0].base == sb ? bins.s[0] : bins.s[1].base == sb ? bins.s[1] : bins.s[2].base == sb ? bins.s[2] : none;
    ct.top.base = tb;
    ct.sec.base = sb;
    slow_finish(t, f, col, ct, bins.total, acgt);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VS_MAX_THREADS, 3) vote_staged_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o,
                                                                        int32_t implied, const TileHdr2 *hdr, const FsTile *fs_tiles) {
    GCB_DYN_SMEM(smem);
    uint64_t *bar = (uint64_t *)(smem + VS_OFF_BAR);
    int *s_nslow = (int *)(smem + VS_OFF_NSLOW);
    int *s_next = (int *)(smem + VS_OFF_NEXT);
    int32_t *s_acc = (int32_t *)(smem + VS_OFF_ACC);
    uint32_t *s_slow = (uint32_t *)(smem + VS_OFF_SLOW);
    FsTile *s_ft = (FsTile *)(smem + VS_OFF_FT);
    VoteRead *s_vr = (VoteRead *)(smem + VS_OFF_VR);
    uint8_t *slab = smem + VS_OFF_SLAB;
#define GCB_LDS32(off) (*(const uint32_t *)(smem + (off)))

    const int tid = (int)threadIdx.x, lane = lane_id(), nthreads = (int)blockDim.x;
    const TileHdr2 h = hdr[blockIdx.x];
    const int nfs = h.nfs;
    if (nfs == 0) return;
    if (tid == 0) {
        tile_barrier_init(bar);
        *s_nslow = 0;
        *s_next = 0;
        const uint32_t vb = 32u * (uint32_t)h.np, fb = 32u * (uint32_t)nfs;
        tile_expect(bar, (uint32_t)h.slab_bytes + vb + fb);
        if (h.slab_bytes > 0) tile_copy(slab, b.payload + h.slab0, (uint32_t)h.slab_bytes, bar);
        tile_copy(s_vr, ws.vote_reads + 2 * (int64_t)h.p0, vb, bar);
        tile_copy(s_ft, fs_tiles + 2 * (int64_t)h.p0, fb, bar);
    }
    for (int i = tid; i < nfs; i += nthreads) s_acc[i] = 0;
    __syncthreads();
    tile_wait(bar, 0);

    TileCtx t;
    t.b = &b; t.r = &r; t.gv = &gv; t.o = &o;
    t.slab = slab; t.vr = s_vr; t.ft = s_ft; t.acc = s_acc;
    t.out0 = r.out_payload + h.out_base0;

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
                if (idx < VS_SLOW_CAP) s_slow[idx] = ((uint32_t)f << 16) | (uint32_t)col;
                else slow_column_general(t, f, col);  // queue full: this lane owns the chunk's words
            }
        }
    }
    __syncthreads();
    // ---- slow columns: one thread per column
    {
        const int n = min(*s_nslow, VS_SLOW_CAP);
        for (int i = tid; i < n; i += nthreads) {
            const uint32_t code = s_slow[i];
            slow_column(t, (int)(code >> 16), (int)(code & 0xFFFFu));
        }
    }
    __syncthreads();
    // ---- per family side: diff, mismatchInc, rollback
    for (int f = tid; f < nfs; f += nthreads) {
        const FsTile ft = s_ft[f];
        if (ft.mode == SIDE_NONE) continue;
        const int acc = s_acc[f];
        const int diff = acc & 0xFFFF, mm = (acc - diff) >> 16;
        if (mm > 5) rollback_record_general(t, f);
        gcb_group_result *gr = r.groups + ft.slot;
        const int side = fs_side(ft);
        gr->diff[side] = diff;
        gr->mismatch_inc[side] = mm;
    }
#undef GCB_LDS32
}

}  // namespace gcb
