// k_vote_pipe.cuh — the vote as a persistent, software-pipelined kernel: the same arithmetic as
// k_vote_tiled.cuh (Pair::computeScore, pair.cpp:88-172, fused with Group::makeConsensus, group.cpp:320-579),
// without its CTA-wide barriers.
//
//   tile_prep_kernel   one CTA per tile (run once per batch, after the output-offset scan): compacts the tile's live
//                      family sides into FsTile entries in global memory, writes the tile header and the absolute
//                      output offsets, and appends the tile to the pipeline's list (or to the generic kernel's).
//   vote_pipe_kernel   one CTA per SM, resident for the whole batch.  Shared memory is a ring of NB stages, each
//                      holding one tile: payload slab, VoteRead table, FsTile table (three bulk asynchronous copies,
//                      cp.async.bulk -> UBLKCP, onto the stage's `full` mbarrier).
//                        warp 0, lane 0   producer: waits for the stage's `empty` mbarrier, writes the stage header,
//                                         issues the copies; runs up to NB tiles ahead of the consumers.
//                        warps 1..15      consumers: take bundles of family sides from the oldest stage that still has
//                                         some (shared-memory counter), vote them exactly like the tiled kernel (sixteen
//                                         columns per lane, hoisted masks for uniform families, 16-bit-lane maxima) and
//                                         queue the slow columns.  The warp that finishes a tile's LAST bundle decides the
//                                         tile's slow columns (one thread per column, the three-bin register histogram of
//                                         the generic kernel) and writes diff / mismatchInc.  Every consumer warp arrives
//                                         on the stage's `empty` barrier when it leaves the tile, so a stage is refilled
//                                         only when nobody can still read it.
//                      No __syncthreads after start-up: a tile's tail overlaps the next tiles' votes.
#pragma once

#include "k_vote_tiled.cuh"

namespace gcb {

constexpr int VP_THREADS = 512;
constexpr int VP_WARPS = VP_THREADS / WARP;
constexpr int VP_MAX_PAIRS = 128;   // pair positions of a pipelined tile
constexpr int VP_MAX_FS = 128;      // family sides of a pipelined tile
constexpr int VP_SLOW_CAP = 256;    // queued slow columns of a tile; more are decided inline by their owner
constexpr int VP_MAX_STAGES = 8;
constexpr int VP_PREP_THREADS = 128;

// per-tile header left by tile_prep_kernel (32 bytes)
struct __align__(16) TileHdr {
    int64_t out_base0;  // first output byte of the tile
    int32_t nfs;        // live family sides (FsTile entries at fs_tiles[2*p0 ..])
    int32_t lanes;      // lanes per family side: the tile's widest record in 16-column chunks
    int32_t common_l;   // l_out of the first family side (mask set computed once per tile)
    int32_t reserved[3];
};

// stage header in shared memory (written by the producer before the stage's `full` barrier completes)
struct __align__(16) StageHdr {
    int64_t out_base0;
    int32_t nfs, lanes, per_bundle, n_bundles, common_l;
    int32_t next_bundle;  // atomic: next bundle to hand out
    int32_t done_bundles; // atomic: bundles finished
    int32_t n_slow;       // atomic: queued slow columns
    int32_t reserved[2];
};
static_assert(sizeof(StageHdr) == 48 && sizeof(TileHdr) == 32, "header sizes");

// shared-memory map: [control][stage 0][stage 1]...
constexpr int VP_OFF_FULL = 0;                          // uint64 full[VP_MAX_STAGES]
constexpr int VP_OFF_EMPTY = 8 * VP_MAX_STAGES;         // uint64 empty[VP_MAX_STAGES]
constexpr int VP_OFF_HDR = 16 * VP_MAX_STAGES;          // StageHdr[VP_MAX_STAGES]
constexpr int VP_OFF_STAGE0 = (VP_OFF_HDR + 48 * VP_MAX_STAGES + 127) & ~127;
// inside a stage
constexpr int VPS_OFF_ACC = 0;                                  // int32[VP_MAX_FS]
constexpr int VPS_OFF_SLOW = VPS_OFF_ACC + 4 * VP_MAX_FS;       // uint32[VP_SLOW_CAP]
constexpr int VPS_OFF_FT = VPS_OFF_SLOW + 4 * VP_SLOW_CAP;      // FsTile[VP_MAX_FS]
constexpr int VPS_OFF_VR = VPS_OFF_FT + 32 * VP_MAX_FS;         // VoteRead[2*VP_MAX_PAIRS]
constexpr int VPS_OFF_SLAB = (VPS_OFF_VR + 32 * VP_MAX_PAIRS + 127) & ~127;
// stage size = VPS_OFF_SLAB + slab_cap + VT_SLAB_SLACK, rounded to 128

// ---- mbarrier helpers for the ring.  Under SIMT-check (one OS thread, fibers) a barrier is a count of completed
// phases and a wait yields to the other fibers until the phase it names is over.
#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ void pipe_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pipe_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pipe_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pipe_expect(uint64_t *bar, uint32_t bytes) { tile_expect(bar, bytes); }
__device__ __forceinline__ void pipe_commit(uint64_t *) {}  // the hardware completes the phase when the bytes have landed
__device__ __forceinline__ void pipe_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GCB_PIPE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GCB_PIPE_DONE;\n"
        "bra GCB_PIPE_WAIT;\n"
        "GCB_PIPE_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void pipe_progress() {}
#else
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
inline void pipe_expect(uint64_t *, uint32_t) {}
inline void pipe_commit(uint64_t *bar) { pipe_arrive(bar); }  // the copies above were synchronous
inline void pipe_wait(uint64_t *bar, uint32_t parity) {
    while (((*bar) & 1u) == parity) ::simt::yield();
}
inline void pipe_progress() { ::simt::st().progress++; }
#endif

// ------------------------------------------------------------------------------------------------
// One CTA per tile: what the tiled kernel's prologue does, once per batch, into global memory.
__global__ void __launch_bounds__(VP_PREP_THREADS) tile_prep_kernel(BatchView b, ResultView r, Workspace ws, int32_t slab_cap, TileHdr *hdr,
                                                                     FsTile *fs_tiles, int32_t *pipe_tiles, int32_t *pipe_count) {
    __shared__ uint32_t s_wsum[VP_PREP_THREADS / WARP];
    __shared__ int s_nofit, s_lmax;
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const TileDir t0 = ws.tile_dir[blockIdx.x], t1 = ws.tile_dir[blockIdx.x + 1];
    const int c0 = t0.c0, c1 = t1.c0;
    if (c0 >= c1) return;
    const int P0 = t0.p0, NP = t1.p0 - t0.p0;
    const int64_t slab_bytes = t1.slab0 - t0.slab0;
    if (NP == 0) return;  // clusters without pairs emit nothing
    if (NP > VP_MAX_PAIRS || slab_bytes > slab_cap) {  // not a tile for the pipeline
        if (tid == 0) {
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = (int32_t)blockIdx.x;
            GCB_COUNT(1, 1);
        }
        return;
    }
    if (tid == 0) {
        s_nofit = 0;
        s_lmax = 1;
    }
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
    const uint32_t mine = (live0 ? 1u : 0u) + (live1 ? 1u : 0u);
    uint32_t incl = mine;
    for (int off = 1; off < WARP; off <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == WARP - 1) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t pre = incl - mine, total = 0;
    for (int w = 0; w < VP_PREP_THREADS / WARP; w++) {
        if (w < warp) pre += s_wsum[w];
        total += s_wsum[w];
    }
    if (tid == 0 && total > (uint32_t)VP_MAX_FS) s_nofit = 1;
    FsTile *ft_out = fs_tiles + 2 * (int64_t)P0;
    int common_l = 0;
    int64_t abs_off[2] = {-1, -1};
    if (live0 || live1) {
        int lneed = 1, fidx = (int)pre;
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
            if ((d.flags & FS_NOFIT) || (orel >> 2) > 0xFFFF || chunks > WARP) s_nofit = 1;
            if (out_base0 + orel + record_bytes(l) > r.out_capacity) {
                raise_error(ws.error_flag, GCB_ERR_CAPACITY);
                ft.mode = SIDE_NONE;  // keeps its place in the table but is never voted
            } else {
                abs_off[side] = out_base0 + orel;
            }
            lneed = max(lneed, min(chunks, WARP));
            if (fidx < VP_MAX_FS) ft_out[fidx] = ft;
            if (fidx == 0) common_l = l;
            fidx++;
        }
        if (lneed > 1) atomicMax(&s_lmax, lneed);
        if (pre == 0 && mine > 0) hdr[blockIdx.x].common_l = common_l;  // the thread that owns family side 0
    }
    __syncthreads();
    if (!s_nofit) {  // the absolute offsets the caller reads (the generic kernel rebases the relative ones itself)
        if (abs_off[0] >= 0) r.groups[P0 + tid].out_off[0] = abs_off[0];
        if (abs_off[1] >= 0) r.groups[P0 + tid].out_off[1] = abs_off[1];
    }
    if (tid == 0) {
        if (s_nofit) {  // the generic kernel takes the tile
            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = (int32_t)blockIdx.x;
            GCB_COUNT(1, 1);
        } else if (total > 0) {
            hdr[blockIdx.x].out_base0 = out_base0;
            hdr[blockIdx.x].nfs = (int32_t)total;
            hdr[blockIdx.x].lanes = s_lmax;
            pipe_tiles[atomicAdd(pipe_count, 1)] = (int32_t)blockIdx.x;
            GCB_COUNT(0, 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// group.cpp:376-525 for one slow column by one thread: the three-bin register histogram of the generic kernel
// (a fourth distinct code falls back to the sixteen-bin local-memory histogram of slow_inline).
GCB_DEV void slow_thread(const TileCtx &t, int f, int col) {
    const FsTile ft = t.ft[f];
    if (col >= (int)ft.len) {
        slow_unvoted(t, f, col);
        return;
    }
    const uint8_t *cb = t.slab + 4 * (int)ft.cbase4;
    const VoteRead *ents = t.vr + ft.ent0;
    const int side = fs_side(ft);
    SparseBins bins;
    bins.init();
    for (int e = 0; e < (int)ft.m; e++) {
        int base, qual, score;
        if (fetch_vote(cb, ents[e], col, side, *t.o, base, qual, score)) bins.add(base, qual, score);
    }
    if (bins.overflow) {
        slow_inline(t, f, col);
        return;
    }
    VoteBin obs[3];
    int nobs = 0;
    uint32_t acgt = 0;
    for (int k = 0; k < 3; k++)
        if (bins.s[k].base >= 0) {
            obs[nobs++] = bins.s[k];
            const int bb = bins.s[k].base;
            if (bb == 1 || bb == 2 || bb == 4 || bb == 8) acgt |= (uint32_t)bins.s[k].maxq << (bb == 1 ? 0 : bb == 2 ? 8 : bb == 4 ? 16 : 24);
        }
    ColumnTop top = column_top(*t.o, obs, nobs, bins.total);
    slow_finish(t, f, col, top, bins.total, acgt);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VP_THREADS, 1) vote_pipe_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o,
                                                                  int32_t implied, int32_t n_stages, int32_t stage_bytes, const TileHdr *hdr,
                                                                  const FsTile *fs_tiles, const int32_t *pipe_tiles, const int32_t *pipe_count) {
    GCB_DYN_SMEM(smem);
    uint64_t *full = (uint64_t *)(smem + VP_OFF_FULL);
    uint64_t *empty = (uint64_t *)(smem + VP_OFF_EMPTY);
    StageHdr *shdr = (StageHdr *)(smem + VP_OFF_HDR);
#define GCB_LDS32(off) (*(const uint32_t *)(smem + (off)))
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const int n_tiles = *pipe_count;
    // this CTA's tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (my_tiles == 0) return;
    // consumer warps are split into one group per stage: group g votes the tiles that land in stage g (k = g, g + NB, ...),
    // so a tile is visited by exactly the warps that can find work in it and the groups run out of phase with each other
    const int wpg = (VP_WARPS - 1) / n_stages;
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) {
            pipe_init(full + s, 1);
            pipe_init(empty + s, wpg);  // every warp of the stage's group arrives once when it leaves the stage's tile
        }
        pipe_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ---- producer: one thread, up to n_stages tiles ahead of the consumers
        if (lane != 0) return;
        // (the directory entries and the header of tile k+1 are fetched while tile k waits for its stage)
        int tile_n = pipe_tiles[blockIdx.x];
        TileDir t0_n = ws.tile_dir[tile_n], t1_n = ws.tile_dir[tile_n + 1];
        TileHdr h_n = hdr[tile_n];
        for (int k = 0; k < my_tiles; k++) {
            const int s = k % n_stages, use = k / n_stages;
            const TileDir t0 = t0_n, t1 = t1_n;
            const TileHdr h = h_n;
            if (k + 1 < my_tiles) {
                tile_n = pipe_tiles[(int64_t)blockIdx.x + (int64_t)(k + 1) * gridDim.x];
                t0_n = ws.tile_dir[tile_n];
                t1_n = ws.tile_dir[tile_n + 1];
                h_n = hdr[tile_n];
            }
            if (use > 0) pipe_wait(empty + s, (uint32_t)((use - 1) & 1));  // the stage's group has left its previous tile
            const int P0 = t0.p0, NP = t1.p0 - t0.p0;
            const uint32_t slab_bytes = (uint32_t)(t1.slab0 - t0.slab0), vr_bytes = 32u * (uint32_t)NP, ft_bytes = 32u * (uint32_t)h.nfs;
            uint8_t *stage = smem + VP_OFF_STAGE0 + (size_t)s * stage_bytes;
            StageHdr sh;
            sh.out_base0 = h.out_base0;
            sh.nfs = h.nfs;
            sh.lanes = h.lanes;
            sh.per_bundle = (int)((32u * ((65535u / (unsigned)h.lanes) + 1u)) >> 16);  // 32 / lanes
            sh.n_bundles = (h.nfs + sh.per_bundle - 1) / sh.per_bundle;
            sh.common_l = h.common_l;
            sh.next_bundle = 0;
            sh.done_bundles = 0;
            sh.n_slow = 0;
            sh.reserved[0] = sh.reserved[1] = 0;
            shdr[s] = sh;
            {
                int4 *acc4 = (int4 *)(stage + VPS_OFF_ACC);
                const int4 zero = {0, 0, 0, 0};
                for (int i = 0; i < (h.nfs + 3) / 4; i++) acc4[i] = zero;
            }
            pipe_expect(full + s, slab_bytes + vr_bytes + ft_bytes);
            if (slab_bytes > 0) tile_copy(stage + VPS_OFF_SLAB, b.payload + t0.slab0, slab_bytes, full + s);
            tile_copy(stage + VPS_OFF_VR, ws.vote_reads + 2 * (int64_t)P0, vr_bytes, full + s);
            tile_copy(stage + VPS_OFF_FT, fs_tiles + 2 * (int64_t)P0, ft_bytes, full + s);
            pipe_commit(full + s);
        }
        return;
    }

    // ---- consumers
    const uint32_t mod4 = 0x01010101u * (uint32_t)(o.moderate_quality & 0xFF);
    const uint32_t sbase = smem_base(smem);
    const int group = (warp - 1) / wpg;
    if (group >= n_stages) return;  // (warps that do not fill a group)
    for (int k = group; k < my_tiles; k += n_stages) {
        const int s = group, use = k / n_stages;
        pipe_wait(full + s, (uint32_t)(use & 1));
        const int stage_off = VP_OFF_STAGE0 + s * stage_bytes;
        uint8_t *stage = smem + stage_off;
        StageHdr *sh = shdr + s;
        int32_t *s_acc = (int32_t *)(stage + VPS_OFF_ACC);
        uint32_t *s_slow = (uint32_t *)(stage + VPS_OFF_SLOW);
        const FsTile *s_ft = (const FsTile *)(stage + VPS_OFF_FT);
        const VoteRead *s_vr = (const VoteRead *)(stage + VPS_OFF_VR);
        const int nfs = sh->nfs, L = sh->lanes, S = sh->per_bundle, nb = sh->n_bundles, common_l = sh->common_l;
        const int64_t out_base0 = sh->out_base0;
        TileCtx t;
        t.b = &b; t.r = &r; t.gv = &gv; t.o = &o;
        t.slab = stage + VPS_OFF_SLAB; t.vr = s_vr; t.ft = s_ft; t.acc = s_acc;
        t.out0 = r.out_payload + out_base0;
        const int sub = (int)(((unsigned)lane * ((65535u / (unsigned)L) + 1u)) >> 16), j = lane - sub * L;
        const int col0 = VT_CHUNK * j;
        const ChunkMasks cm_common = make_masks(common_l, common_l, col0);
        bool last = false;  // this warp finished the tile's last bundle
        for (;;) {
            int bundle = 0;
            if (lane == 0) bundle = atomicAdd(&sh->next_bundle, 1);
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
            const int cb = stage_off + VPS_OFF_SLAB + 4 * (int)ft.cbase4;  // byte offsets into the CTA's shared memory
            const int ento = stage_off + VPS_OFF_VR + 16 * (int)ft.ent0;
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
            uint32_t mo[4] = {0u, 0u, 0u, 0u}, me[4] = {0u, 0u, 0u, 0u}, dis0 = 0u, dis1 = 0u;
            if (ft.flags & FS_UNIFORM) {
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
                    for (int kk = 0; kk < 4; kk++) {
                        mo[kk] = __vmaxu2(mo[kk], q[kk]);
                        me[kk] = __vmaxu2(me[kk], q[kk] << 8);
                    }
                    dis0 |= (be0 ^ tbe0) & vm0;
                    dis1 |= (be1 ^ tbe1) & vm1;
                    if (v.ov_len > 0) {  // (subtractions only: see the ptxas note in k_vote_tiled.cuh)
                        const int x = (int)v.ov_own - rp0;
                        const int y = x - (int)v.ov_mate;
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
            if (mine) {
                uint32_t oq[4];
                uint32_t slow0 = 0u, slow1 = 0u;
                if (ft.mode == SIDE_COPY) {  // group.cpp:73-77: the record itself
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) oq[kk] = col0 + 4 * kk < qbytes ? GCB_LDS32(trec + col0 + 4 * kk) : 0u;
                } else {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) oq[kk] = prmt(mo[kk], me[kk], 0x3715u) & cm.vb[kk];
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
                for (int kk = 0; kk < 4; kk++)
                    if (col0 + 4 * kk < qbytes) *(uint32_t *)(out + col0 + 4 * kk) = oq[kk] & cm.rb[kk];
                if (8 * j < sbytes) *(uint32_t *)(out + qbytes + 8 * j) = bswap32(tbe0 & cm.kn0);
                if (8 * j + 4 < sbytes) *(uint32_t *)(out + qbytes + 8 * j + 4) = bswap32(tbe1 & cm.kn1);
                for (int wsel = 0; wsel < 2; wsel++) {
                    uint32_t sm = wsel ? slow1 : slow0;
                    while (sm != 0u) {
                        const int kk = __clz((int)sm) >> 2;
                        sm &= ~(0xF0000000u >> (4 * kk));
                        const int col = col0 + 8 * wsel + kk;
                        GCB_COUNT(3, 1);
                        const int idx = atomicAdd(&sh->n_slow, 1);
                        if (idx < VP_SLOW_CAP) s_slow[idx] = ((uint32_t)f << 16) | (uint32_t)col;
                        else slow_thread(t, f, col);  // queue full: this lane owns the chunk's words (diff goes to s_acc below)
                    }
                }
            }
            // the bundle is done: its stores and queue entries must be visible to whoever ends the tile
            __threadfence_block();
            __syncwarp();
            int done = 0;
            if (lane == 0) done = atomicAdd(&sh->done_bundles, 1) + 1;
            done = __shfl_sync(FULL, done, 0);
            pipe_progress();
            if (done == nb) {
                last = true;
                break;
            }
        }
        if (!last) {  // nothing left for this warp in the tile: it will not touch the stage again
            __syncwarp();
            if (lane == 0) pipe_arrive(empty + s);
            continue;
        }
        // ---- this warp ends the tile: slow columns (one thread each), diff / mismatchInc / rollback, release the stage
        __threadfence_block();
        const int n = min(sh->n_slow, VP_SLOW_CAP);
        for (int i = lane; i < n; i += WARP) slow_thread(t, (int)(s_slow[i] >> 16), (int)(s_slow[i] & 0xFFFFu));
        __syncwarp();
        for (int f = lane; f < nfs; f += WARP) {
            const FsTile ft = s_ft[f];
            if (ft.mode == SIDE_NONE) continue;
            const int acc = s_acc[f];
            const int diff = acc & 0xFFFF, mm = (acc - diff) >> 16;
            if (mm > 5) rollback_record(t, f);
            gcb_group_result *gr = r.groups + ft.slot;
            const int side = fs_side(ft);
            gr->diff[side] = diff;
            gr->mismatch_inc[side] = mm;
        }
        __syncwarp();
        if (lane == 0) pipe_arrive(empty + s);
    }
#undef GCB_LDS32
}

}  // namespace gcb
