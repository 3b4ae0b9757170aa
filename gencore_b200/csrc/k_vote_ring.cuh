// k_vote_ring.cuh — the fast half of the split vote (k_vote_split.cuh) as ONE persistent CTA per SM over a ring of
// staged tiles: Pair::computeScore (pair.cpp:88-172) fused with Group::makeConsensus (group.cpp:320-579).
//
//   warp 0, lane 0    producer.  Walks this CTA's tiles (blockIdx.x, + gridDim.x, ...), waits for the stage's `empty`
//                     mbarrier, writes the 64-byte stage header and issues three bulk asynchronous copies
//                     (cp.async.bulk -> UBLKCP) onto the stage's `full` mbarrier: compact family-side list, VoteRead table,
//                     payload slab.  Runs up to n_stages tiles ahead; the last stage it fills carries nfs < 0.
//   warps 1..15       consumers.  Every warp visits every tile in order: waits for `full`, takes bundles of family sides
//                     from the stage's counter until none is left, arrives on `empty`.  A tile has fewer bundles than the
//                     CTA has warps, so the warps spread over the tiles in flight; nobody waits for a tile's last column:
//                     the slow columns are queued for slow_columns_kernel (k_vote_split.cuh).
//   a bundle          lanes_per_side lanes per family side, sixteen columns per lane: the branch-free uniform loop of
//                     vote_fast_kernel (raw-order XOR residues, VIMNMX3.U16x2 over two reads, mate alignment once per
//                     bundle).  Slow columns of the bundle are emitted COOPERATIVELY: the lanes list their columns in a
//                     small per-warp table, one 64-bit atomic reserves all records, then eight lanes per column write the
//                     column's entries (one lane per read), in rounds of at most VR_ITEMS columns.
//   queue overflow    the tile is handed to the generic kernel (score_vote_kernel), which runs last and rewrites all of the
//                     tile's records from the payload.
#pragma once

#include "k_vote_split.cuh"

namespace gcb {

constexpr int VR_MAX_THREADS = 768;           // the kernel is instantiated for 512 and 768 threads (128 / 85 registers)
constexpr int VR_WARPS = VR_MAX_THREADS / WARP;
constexpr int VR_MAX_STAGES = 8;   // tiles in flight (barrier pairs and stage headers); their bytes come from one arena
constexpr int VR_GUARD = 4608;     // never allocated, after the arena: the branch-free read loop may read a VoteRead table or a
                                   // slab up to 257 entries / 64 bytes past its end (values unused)
constexpr int VR_ITEMS = 64;      // sparse slow columns of one bundle that are emitted cooperatively
constexpr int VR_GROUP = 8;       // lanes per slow column in the cooperative emission

struct __align__(16) RingStage {  // shared memory, written by the producer before the stage's `full` barrier completes
    int64_t out_base0;
    int32_t nfs;          // live family sides; < 0: no more tiles
    int32_t lanes, per_bundle, n_bundles, common_l;
    int32_t p0, tile;
    int32_t next_bundle;  // atomic: next bundle to hand out
    int32_t handed_over;  // atomic: the tile went to the generic kernel (queue overflow)
    int32_t ft_off, vr_off, slab_off;  // where the tile's family-side list, VoteRead table and payload slab lie (shared-memory offsets)
    int32_t first_ticket;              // bundles of this CTA's earlier tiles, modulo the number of consumer warps
    int32_t pad[1];
};
static_assert(sizeof(RingStage) == 64, "stage header size");

// shared-memory map: [barriers][stage headers][per-warp slow-column tables][stage 0][stage 1]...
constexpr int VR_OFF_FULL = 0;                                   // uint64[VR_MAX_STAGES]
constexpr int VR_OFF_EMPTY = 8 * VR_MAX_STAGES;                  // uint64[VR_MAX_STAGES]
constexpr int VR_OFF_HDR = 128;                                  // RingStage[VR_MAX_STAGES]
constexpr int VR_OFF_ITEMS = VR_OFF_HDR + 64 * VR_MAX_STAGES;    // per warp: uint16 codes[VR_ITEMS] (family side in the bundle << 9 | column)
constexpr int VR_ITEM_BYTES = 2 * VR_ITEMS;
constexpr int VR_POOL_RECS = 64, VR_POOL_WORDS = 64 * 20;  // queue space a warp reserves at a time
constexpr int VR_OFF_HCACHE = (VR_OFF_ITEMS + VR_ITEM_BYTES * VR_WARPS + 15) & ~15;  // TileHdr2[32]: the producer's next tiles
constexpr int VR_OFF_START = VR_OFF_HCACHE + 48 * WARP;          // uint32[VR_MAX_STAGES]: the producer's allocation starts
constexpr int VR_OFF_ARENA = (VR_OFF_START + 4 * VR_MAX_STAGES + 127) & ~127;
// a tile's allocation: [FsTile list][VoteRead table][slab + slack], each part rounded to 128 bytes
static_assert(16 * VR_MAX_STAGES <= VR_OFF_HDR && VR_OFF_ARENA % 128 == 0, "ring layout");
GCB_HD uint32_t ring_round128(uint32_t v) { return (v + 127u) & ~127u; }

// a wait that lets the hardware suspend the thread between polls (a spinning warp takes issue slots from the warps that vote)
#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ void pipe_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns) {
    const uint32_t a = smem_u32(bar);
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity), "r"(ns)
            : "memory");
        if (done) return;
    }
}
#else
inline void pipe_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t) { pipe_wait(bar, parity); }
#endif

template <int NT, int NU>
__global__ void __launch_bounds__(NT, 1) vote_ring_kernel(BatchView b, ResultView r, Workspace ws, int32_t moderate_quality, int32_t implied,
                                                                  const TileHdr2 *hdr, const FsTile *fs_tiles, SlowQueues sq, int32_t n_tiles,
                                                                  int32_t max_stages, int32_t arena_bytes, const int32_t *max_need,
                                                                  int32_t ablate) {
    // `ablate` (profiling only, 0 in production; results are wrong otherwise): 1 = no slow-column emission, 2 = no read loop,
    // 4 = no record stores, 8 = no bundle work at all
    GCB_DYN_SMEM(smem);
    uint64_t *full = (uint64_t *)(smem + VR_OFF_FULL);
    uint64_t *empty = (uint64_t *)(smem + VR_OFF_EMPTY);
    RingStage *shdr = (RingStage *)(smem + VR_OFF_HDR);
#define GCB_LDS32(off) (*(const uint32_t *)(smem + (off)))
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = tid >> 5;
    // the arena is cut into equal stages that hold the batch's largest tile (tile_prep2_kernel measured it): as many tiles in
    // flight as fit
    const int32_t stage_bytes = (int32_t)ring_round128((uint32_t)max(*max_need, 128));
    const int n_stages = min(max_stages, max(arena_bytes / stage_bytes, 1));
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) {
            pipe_init(full + s, 1);
            pipe_init(empty + s, NT / WARP - 1);  // every consumer warp arrives once when it leaves the stage's tile
        }
        pipe_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ---- producer: lane 0 fills the stages, up to n_stages tiles ahead of the consumers; the whole warp fetches the
        // headers of this CTA's next 32 tiles at once, so that no tile waits for a header on its way from global memory
        TileHdr2 *hcache = (TileHdr2 *)(smem + VR_OFF_HCACHE);
        int k = 0;
        // stage k % n_stages once every consumer has left the tile it held before
        auto allocate = [&]() -> uint32_t {
            const int s = k % n_stages, use = k / n_stages;
            if (use > 0) pipe_wait_backoff(empty + s, (uint32_t)((use - 1) & 1), 2000u);
            return (uint32_t)s * (uint32_t)stage_bytes;
        };
        for (int64_t base = (int64_t)blockIdx.x; base < n_tiles; base += (int64_t)WARP * gridDim.x) {
            const int64_t mine_t = base + (int64_t)lane * gridDim.x;
            if (mine_t < n_tiles) hcache[lane] = hdr[mine_t];
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < WARP; i++) {
                    const int64_t t = base + (int64_t)i * gridDim.x;
                    if (t >= n_tiles) break;
                    const TileHdr2 cur = hcache[i];
                    if (cur.nfs <= 0) continue;  // nothing for this kernel here (empty tile, or the generic kernel has it)
                    const uint32_t slab_bytes = (uint32_t)cur.slab_bytes, vr_bytes = 32u * (uint32_t)cur.np, ft_bytes = 32u * (uint32_t)cur.nfs;
                    const uint32_t at = allocate();
                    const int s = k % n_stages;
                    RingStage sh;
                    sh.out_base0 = cur.out_base0;
                    sh.nfs = cur.nfs;
                    sh.lanes = cur.lanes; sh.per_bundle = cur.per_bundle; sh.n_bundles = cur.n_bundles; sh.common_l = cur.common_l;
                    sh.p0 = cur.p0; sh.tile = (int32_t)t;
                    sh.next_bundle = 0;
                    sh.handed_over = 0;
                    sh.ft_off = VR_OFF_ARENA + (int32_t)at;
                    sh.vr_off = sh.ft_off + (int32_t)ring_round128(ft_bytes);
                    sh.slab_off = sh.vr_off + (int32_t)ring_round128(vr_bytes);
                    sh.first_ticket = 0;
                    sh.pad[0] = 0;
                    shdr[s] = sh;
                    pipe_expect(full + s, slab_bytes + vr_bytes + ft_bytes);
                    if (slab_bytes > 0) tile_copy(smem + sh.slab_off, b.payload + cur.slab0, slab_bytes, full + s);
                    tile_copy(smem + sh.vr_off, ws.vote_reads + 2 * (int64_t)cur.p0, vr_bytes, full + s);
                    tile_copy(smem + sh.ft_off, fs_tiles + 2 * (int64_t)cur.p0, ft_bytes, full + s);
                    pipe_commit(full + s);
                    k++;
                }
            }
            __syncwarp();
        }
        if (lane == 0) {  // the end marker: the phase completes with this arrival alone
            allocate();
            const int s = k % n_stages;
            RingStage sh;
            sh.out_base0 = 0;
            sh.nfs = -1;
            sh.lanes = 1; sh.per_bundle = 32; sh.n_bundles = 0; sh.common_l = 0;
            sh.p0 = 0; sh.tile = 0;
            sh.next_bundle = 0;
            sh.handed_over = 0;
            sh.ft_off = sh.vr_off = sh.slab_off = VR_OFF_ARENA;
            sh.first_ticket = 0;
            sh.pad[0] = 0;
            shdr[s] = sh;
            pipe_expect(full + s, 0u);
            pipe_commit(full + s);
        }
        return;
    }

    // ---- consumers
    const uint32_t mod4 = 0x01010101u * (uint32_t)(moderate_quality & 0xFF);
    const uint32_t sbase = smem_base(smem);
    uint16_t *s_item = (uint16_t *)(smem + VR_OFF_ITEMS + VR_ITEM_BYTES * warp);
    // slow-column records go to the queue of this CTA; every warp keeps a pool of reserved records / words in registers
    const int qi = (int)(blockIdx.x % VQ_NQ);
    uint32_t *q_words = sq.words + (size_t)qi * sq.cap_words;
    uint32_t *q_index = sq.index + (size_t)qi * sq.cap_recs;
    uint32_t pool_r = 0u, pool_re = 0u, pool_w = 0u, pool_we = 0u;
    int s = 0;
    uint32_t par = 0u;
    for (;;) {
        pipe_wait_backoff(full + s, par, 1000u);
        RingStage *sh = shdr + s;
        const int nfs = sh->nfs;
        if (nfs < 0) break;
        // a lane owns NU units of sixteen columns; a family side takes L lanes, a bundle 32 / L family sides
        const int L = (sh->lanes + NU - 1) / NU;
        const int S = (int)((32u * ((65535u / (unsigned)L) + 1u)) >> 16);
        const int nb = NU == 1 ? sh->n_bundles : (int)(((unsigned)(nfs + S - 1) * ((65535u / (unsigned)S) + 1u)) >> 16);
        int bundle = nb;
        if (lane == 0 && *(volatile int32_t *)&sh->next_bundle < nb) bundle = atomicAdd(&sh->next_bundle, 1);  // (no atomic on a drained tile)
        bundle = __shfl_sync(FULL, bundle, 0);
        if (bundle < nb) {
            const int off_slab = sh->slab_off, off_vr = sh->vr_off;
            const FsTile *s_ft = (const FsTile *)(smem + sh->ft_off);
            const VoteRead *s_vr = (const VoteRead *)(smem + off_vr);
            const int64_t out_base0 = sh->out_base0;
            uint8_t *out0 = r.out_payload + out_base0;
            const int tile = sh->tile;
            const int sub = (int)(((unsigned)lane * ((65535u / (unsigned)L) + 1u)) >> 16), j = lane - sub * L;  // lane / L, lane % L
            const int col0 = VT_CHUNK * NU * j;
            const int common_l = sh->common_l;  // the masks of the tile's usual record length are computed once
            ChunkMasks cm_common[NU];
#pragma unroll
            for (int u = 0; u < NU; u++) cm_common[u] = make_masks(common_l, common_l, col0 + VT_CHUNK * u);
            do {
                if (ablate & 8) goto next_bundle;
                {
                const int f = bundle * S + sub;
                FsTile ft;
                ft.ent0 = 0; ft.m = 0; ft.l_out = 0; ft.len = 0; ft.tmpl_k = 0; ft.mode = SIDE_NONE; ft.flags = 0; ft.cbase4 = 0; ft.out4 = 0;
                if (sub < S && f < nfs) ft = s_ft[f];
                const int l_out = ft.l_out, len = ft.len;
                const int qbytes = GCB_ALIGN4(l_out), sbytes = GCB_ALIGN4((l_out + 1) >> 1);
                const bool mine = ft.mode != SIDE_NONE && col0 < max(qbytes, 2 * sbytes);  // this lane owns words of the record
                const int m = mine && ft.mode != SIDE_COPY ? (int)ft.m : 0;
                const int mmax = (ablate & 2) ? 0 : __reduce_max_sync(FULL, m);
                const int cb = off_slab + 4 * (int)ft.cbase4;  // byte offsets into the CTA's shared memory
                const int ento = off_vr + 16 * (int)ft.ent0;
                VoteRead tv = {0, 0, 0, 0, 0, 0, 0, 0};
                uint32_t tbe[2 * NU];
#pragma unroll
                for (int w = 0; w < 2 * NU; w++) tbe[w] = 0u;
                int trec = cb;
                const int sb0 = 8 * NU * j;  // the lane's first byte of a record's packed bases
                if (mine) {
                    tv = s_vr[ft.ent0 + ft.tmpl_k];
                    trec = cb + 4 * (int)tv.own_off4;
#pragma unroll
                    for (int w = 0; w < 2 * NU; w++)
                        if (sb0 + 4 * w < sbytes) tbe[w] = bswap32(GCB_LDS32(trec + qbytes + sb0 + 4 * w));
                }
                ChunkMasks cm[NU];
#pragma unroll
                for (int u = 0; u < NU; u++) cm[u] = cm_common[u];
                if (l_out != common_l || len != l_out) {
#pragma unroll
                    for (int u = 0; u < NU; u++) cm[u] = make_masks(l_out, len, col0 + VT_CHUNK * u);
                }
                if (mine && j == 0 && ft.mode != SIDE_COPY) GCB_COUNT((ft.flags & FS_UNIFORM) ? 4 : 5, 1);
                if (mine && j == 0) {  // slow_columns_kernel adds to these
                    gcb_group_result *gr = r.groups + ft.slot;
                    const int sd = (ft.flags & FS_SIDE1) ? 1 : 0;
                    gr->diff[sd] = 0;
                    gr->mismatch_inc[sd] = 0;
                }
                uint32_t mo[4 * NU], me[4 * NU], dis[2 * NU];
#pragma unroll
                for (int w = 0; w < 4 * NU; w++) mo[w] = me[w] = 0u;
#pragma unroll
                for (int w = 0; w < 2 * NU; w++) dis[w] = 0u;
                if (ft.flags & FS_UNIFORM) {  // (see vote_fast_kernel)
                    uint32_t om[2 * NU];
                    bool has_ov = false;
                    const int x0 = (int)tv.ov_own - col0, y0 = x0 - (int)tv.ov_mate;
#pragma unroll
                    for (int u = 0; u < NU; u++) {
                        const int x = x0 - VT_CHUNK * u, y = y0 - VT_CHUNK * u;  // first column of the unit inside the window / with a mate index >= 0
                        const int oa = max(max(0, x), y), oz = min(min(cm[u].nvote, x + (int)tv.ov_len), y + (int)tv.mate_l);
                        const bool hu = tv.ov_len > 0 && oz > oa;
                        om[2 * u] = hu ? nib_range(oa, oz) : 0u;
                        om[2 * u + 1] = hu ? nib_range(oa - 8, oz - 8) : 0u;
                        has_ov = has_ov || hu;
                    }
                    const int ms = 0 - y0, mw0 = ms >> 3;  // the lane's first mate column: word mw0 of the mate's bases, nibble ms & 7
                    const unsigned msh = (unsigned)(ms & 7) * 4u;
                    const uint32_t qbase = sbase + (uint32_t)(cb + col0), sdelta = (uint32_t)(qbytes - col0 + sb0);
                    const uint32_t mdelta = has_ov ? (uint32_t)(4 * ((int)tv.mate_off4 - (int)tv.own_off4) + GCB_ALIGN4(tv.mate_l) + 4 * mw0 - col0) : 0u;
                    const uint32_t xt = tv.own_off4;
                    const uint32_t qt = qbase + (xt << 2);
                    uint32_t t[2 * NU], a0[2 * NU + 1], d[2 * NU], da[2 * NU + 1];
#pragma unroll
                    for (int w = 0; w < 2 * NU; w++) {
                        t[w] = lds32r(qt + sdelta + 4 * w);
                        d[w] = 0u;
                    }
#pragma unroll
                    for (int w = 0; w < 2 * NU + 1; w++) {
                        a0[w] = lds32r(qt + mdelta + 4 * w);
                        da[w] = 0u;
                    }
                    uint32_t ea = sbase + (uint32_t)ento;
                    for (int e = 0; e < mmax; e += 2, ea += 32) {
                        uint32_t xa = lds16<0>(ea), xb = lds16<16>(ea);
                        xa = (e < m && xa != VR_NO_VOTE) ? xa : xt;
                        xb = (e + 1 < m && xb != VR_NO_VOTE) ? xb : xt;
                        const uint32_t qa = qbase + (xa << 2), qb = qbase + (xb << 2);
#pragma unroll
                        for (int w = 0; w < 4 * NU; w++) {
                            const uint32_t va = lds32r(qa + 4 * w), vb = lds32r(qb + 4 * w);
                            mo[w] = __vimax3_u16x2(mo[w], va, vb);
                            me[w] = __vimax3_u16x2(me[w], va << 8, vb << 8);
                        }
#pragma unroll
                        for (int w = 0; w < 2 * NU; w++) d[w] |= (lds32r(qa + sdelta + 4 * w) ^ t[w]) | (lds32r(qb + sdelta + 4 * w) ^ t[w]);
#pragma unroll
                        for (int w = 0; w < 2 * NU + 1; w++) da[w] |= (lds32r(qa + mdelta + 4 * w) ^ a0[w]) | (lds32r(qb + mdelta + 4 * w) ^ a0[w]);
                    }
#pragma unroll
                    for (int w = 0; w < 2 * NU; w++) dis[w] = bswap32(d[w]);
                    if (has_ov) {  // pair.cpp:133-170: a base that differs from its mate's is never a fast column
#pragma unroll
                        for (int w = 0; w < 2 * NU + 1; w++) {
                            da[w] = bswap32(da[w]);
                            a0[w] = bswap32(a0[w]);
                        }
#pragma unroll
                        for (int w = 0; w < 2 * NU; w++)
                            dis[w] |= (__funnelshift_l(da[w + 1], da[w], msh) | (bswap32(t[w]) ^ __funnelshift_l(a0[w + 1], a0[w], msh))) & om[w];
                    }
                } else if (m > 0) {
                    for (int e = 0; e < m; e++) {
                        const VoteRead v = s_vr[ft.ent0 + e];
                        if (v.own_off4 == VR_NO_VOTE || v.own_l == 0) continue;
                        const uint8_t *rec = smem + cb + 4 * (int)v.own_off4;
                        const int rq = GCB_ALIGN4(v.own_l);
#pragma unroll
                        for (int u = 0; u < NU; u++) {
                            const int rp0 = col0 + VT_CHUNK * u + v.shift;
                            const int a = max(0, 0 - rp0), z = min(cm[u].nvote, (int)v.own_l - rp0);
                            if (z <= a) continue;
                            uint32_t q[4], be0, be1;
                            fetch16q(rec, rq, rp0, q);
                            fetch16b(rec + rq, GCB_ALIGN4((v.own_l + 1) >> 1), rp0, be0, be1);
                            const uint32_t vm0 = nib_range(a, z), vm1 = nib_range(a - 8, z - 8);
                            q[0] &= bytes_lo(vm0); q[1] &= bytes_hi(vm0); q[2] &= bytes_lo(vm1); q[3] &= bytes_hi(vm1);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                mo[4 * u + kk] = __vmaxu2(mo[4 * u + kk], q[kk]);
                                me[4 * u + kk] = __vmaxu2(me[4 * u + kk], q[kk] << 8);
                            }
                            dis[2 * u] |= (be0 ^ tbe[2 * u]) & vm0;
                            dis[2 * u + 1] |= (be1 ^ tbe[2 * u + 1]) & vm1;
                            if (v.ov_len > 0) {  // (subtractions only: see the ptxas note in k_vote_tiled.cuh)
                                const int x = (int)v.ov_own - rp0;
                                const int y = x - (int)v.ov_mate;
                                const int oa = max(max(a, x), y);
                                const int oz = min(min(z, x + (int)v.ov_len), y + (int)v.mate_l);
                                if (oz > oa) {
                                    const uint8_t *mrec = smem + cb + 4 * (int)v.mate_off4;
                                    uint32_t mb0, mb1;
                                    fetch16b(mrec + GCB_ALIGN4(v.mate_l), GCB_ALIGN4((v.mate_l + 1) >> 1), 0 - y, mb0, mb1);
                                    dis[2 * u] |= (be0 ^ mb0) & nib_range(oa, oz);
                                    dis[2 * u + 1] |= (be1 ^ mb1) & nib_range(oa - 8, oz - 8);
                                }
                            }
                        }
                    }
                }
                // ---- what the record gets: qualities = the maxima (fast columns), bases = the template's
                uint32_t slow[2 * NU];  // one bit per slow column (the low bit of its big-endian nibble)
#pragma unroll
                for (int w = 0; w < 2 * NU; w++) slow[w] = 0u;
                if (mine) {
                    uint8_t *out = out0 + 4 * (int64_t)ft.out4;
#pragma unroll
                    for (int u = 0; u < NU; u++) {
                        const int cu = col0 + VT_CHUNK * u;
                        uint32_t oq[4];
                        if (ft.mode == SIDE_COPY) {  // group.cpp:73-77: the record itself
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) oq[kk] = cu + 4 * kk < qbytes ? GCB_LDS32(trec + cu + 4 * kk) : 0u;
                        } else {
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) oq[kk] = prmt(mo[4 * u + kk], me[4 * u + kk], 0x3715u) & cm[u].vb[kk];  // (the hoisted loop read whole words)
                            GCB_COUNT(2, cm[u].nvote);
                            if (implied && len == l_out) {
                                const uint32_t lowq0 = nibs_of_flags(bytes_ge_flags(oq[0], mod4) ^ 0x80808080u, bytes_ge_flags(oq[1], mod4) ^ 0x80808080u);
                                const uint32_t lowq1 = nibs_of_flags(bytes_ge_flags(oq[2], mod4) ^ 0x80808080u, bytes_ge_flags(oq[3], mod4) ^ 0x80808080u);
                                uint32_t s0 = (dis[2 * u] | lowq0) & cm[u].vn0, s1 = (dis[2 * u + 1] | lowq1) & cm[u].vn1;
                                s0 |= s0 >> 1; s0 |= s0 >> 2; s0 &= 0x11111111u;  // any differing bit marks the column
                                s1 |= s1 >> 1; s1 |= s1 >> 2; s1 &= 0x11111111u;
                                slow[2 * u] = s0;
                                slow[2 * u + 1] = s1;
                            } else {  // without `implied`, or with columns that are not voted, every column of the record is slow
                                slow[2 * u] = nibs_of_bytes(cm[u].rb[0], cm[u].rb[1]) & 0x11111111u;
                                slow[2 * u + 1] = nibs_of_bytes(cm[u].rb[2], cm[u].rb[3]) & 0x11111111u;
                            }
                        }
                        if (!(ablate & 4)) {
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)
                                if (cu + 4 * kk < qbytes) *(uint32_t *)(out + cu + 4 * kk) = oq[kk] & cm[u].rb[kk];
                            if (sb0 + 8 * u < sbytes) *(uint32_t *)(out + qbytes + sb0 + 8 * u) = bswap32(tbe[2 * u] & cm[u].kn0);
                            if (sb0 + 8 * u + 4 < sbytes) *(uint32_t *)(out + qbytes + sb0 + 8 * u + 4) = bswap32(tbe[2 * u + 1] & cm[u].kn1);
                        }
                    }
                }
                // ---- slow columns: records of one size per bundle, taken from the warp's own pool of reserved queue space
                int nslow = 0;
#pragma unroll
                for (int w = 0; w < 2 * NU; w++) nslow += __popc(slow[w]);
                if (ablate & 1) nslow = 0;
                const uint32_t T = (uint32_t)__reduce_add_sync(FULL, nslow);
                if (T > 0) {
                    GCB_COUNT(3, nslow);
                    const uint32_t rw = slow_rec_words(mmax), W = T * rw;
                    if (pool_r + T > pool_re || pool_w + W > pool_we) {
                        // a new pool (one 64-bit atomic: records << 32 | words); what is left of the old one stays unused
                        for (uint32_t i = pool_r + (uint32_t)lane; i < pool_re; i += WARP) q_index[i] = VQ_INVALID;
                        const uint32_t need_r = T > (uint32_t)VR_POOL_RECS ? T : (uint32_t)VR_POOL_RECS;
                        const uint32_t need_w = W > (uint32_t)VR_POOL_WORDS ? W : (uint32_t)VR_POOL_WORDS;
                        unsigned long long base64 = 0ull;
                        if (lane == 0) base64 = atomicAdd(sq.count + qi, ((unsigned long long)need_r << 32) | need_w);
                        base64 = __shfl_sync(FULL, base64, 0);
                        const uint32_t r0 = (uint32_t)(base64 >> 32), w0 = (uint32_t)base64;
                        if ((unsigned long long)r0 + need_r <= sq.cap_recs && (unsigned long long)w0 + need_w <= sq.cap_words) {
                            pool_r = r0; pool_re = r0 + need_r;
                            pool_w = w0; pool_we = w0 + need_w;
                        } else {  // the queue is full: the reserved index entries are marked unused, the pool stays empty
                            for (uint32_t i = r0 + (uint32_t)lane; i < r0 + need_r && i < sq.cap_recs; i += WARP) q_index[i] = VQ_INVALID;
                            pool_r = pool_re = pool_w = pool_we = 0u;
                        }
                    }
                    if (pool_r + T > pool_re) {
                        // no queue space: the generic kernel redoes the whole tile from the payload (it runs after
                        // slow_columns_kernel and vote_rollback_kernel)
                        if (lane == 0 && atomicExch(&sh->handed_over, 1) == 0) {
                            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = ~tile;
                            GCB_COUNT(1, 1);
                        }
                    } else {
                        // rounds of at most VR_ITEMS columns: every lane lists one of its columns per ballot, then eight lanes per
                        // column write the column's entries, one lane per read
                        unsigned long long sm64[NU];
#pragma unroll
                        for (int u = 0; u < NU; u++) sm64[u] = ((unsigned long long)slow[2 * u] << 32) | slow[2 * u + 1];
                        uint32_t emitted = 0u;
                        while (emitted < T) {
                            uint32_t listed = 0u;
                            while (listed + WARP <= (uint32_t)VR_ITEMS) {
                                bool has = false;
                                int u_pick = 0;
#pragma unroll
                                for (int u = NU - 1; u >= 0; u--)
                                    if (sm64[u] != 0ull) {
                                        has = true;
                                        u_pick = u;
                                    }
                                const unsigned bal = __ballot_sync(FULL, has);
                                if (bal == 0u) break;
                                if (has) {
                                    unsigned long long v = sm64[0];
#pragma unroll
                                    for (int u = 1; u < NU; u++)
                                        if (u_pick == u) v = sm64[u];
                                    const int kk = __clzll((long long)v) >> 2;
                                    v &= ~(0xF000000000000000ull >> (4 * kk));
#pragma unroll
                                    for (int u = 0; u < NU; u++)
                                        if (u_pick == u) sm64[u] = v;
                                    s_item[listed + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)((sub << 9) | (col0 + VT_CHUNK * u_pick + kk));
                                }
                                listed += __popc(bal);
                            }
                            __syncwarp();
                            const int g = lane / VR_GROUP, gl = lane % VR_GROUP;
                            for (uint32_t i = (uint32_t)g; i < listed; i += WARP / VR_GROUP) {
                                const uint32_t code = s_item[i], wofs = pool_w + (emitted + i) * rw;
                                const int col = (int)(code & 511u), fi = bundle * S + (int)(code >> 9);
                                const FsTile fti = s_ft[fi];
                                const uint8_t *cbp = smem + off_slab + 4 * (int)fti.cbase4;
                                const VoteRead *ents = s_vr + fti.ent0;
                                uint32_t *rec = q_words + wofs;
                                const int mi = (int)fti.m;
                                if (gl == 0) {
                                    q_index[pool_r + emitted + i] = wofs;
                                    slow_write_header(rec, fti, col, out_base0 + 4 * (int64_t)fti.out4);
                                }
                                if ((fti.flags & FS_UNIFORM) && col < (int)fti.len) {
                                    // the column's place in pair.cpp:121-170 is the same for every read of a uniform family
                                    const VoteRead tvi = ents[fti.tmpl_k];
                                    const bool info = tvi.ov_len != VR_NO_OVERLAP_INFO;
                                    const int kq = col - (int)tvi.ov_own, mp = (int)tvi.ov_mate + kq;
                                    const bool inwin = info && kq >= 0 && kq < (int)tvi.ov_len;
                                    const bool mvalid = inwin && mp >= 0 && mp < (int)tvi.mate_l;
                                    const uint32_t st = !info ? SE_NO_INFO : !inwin ? SE_PLAIN : mvalid ? SE_MATE : SE_NO_MATE_BASE;
                                    const int soff = GCB_ALIGN4(fti.l_out) + (col >> 1), nsh = (col & 1) ? 0 : 4;
                                    const int mrel = mvalid ? 4 * ((int)tvi.mate_off4 - (int)tvi.own_off4) : 0, mpi = mvalid ? mp : 0;
                                    const int mqoff = mrel + mpi, msoff = mrel + (mvalid ? GCB_ALIGN4(tvi.mate_l) : 0) + (mpi >> 1), mnsh = (mpi & 1) ? 0 : 4;
                                    for (int e = gl; e < mi; e += VR_GROUP) {
                                        const uint32_t xo = ents[e].own_off4;
                                        uint32_t ent = 0u;
                                        if (xo != VR_NO_VOTE) {
                                            const uint8_t *p = cbp + 4 * (int)xo;
                                            const uint32_t ql = p[col], base = ((uint32_t)p[soff] >> nsh) & 0xFu;
                                            const uint32_t mql = mvalid ? p[mqoff] : 0u, mbase = mvalid ? (((uint32_t)p[msoff] >> mnsh) & 0xFu) : 0u;
                                            ent = ql | (mql << 8) | (base << 16) | (mbase << 20) | (st << 24) | SE_VOTES;
                                        }
                                        rec[SR_HDR_WORDS + e] = ent;
                                    }
                                } else {
                                    for (int e = gl; e < mi; e += VR_GROUP) rec[SR_HDR_WORDS + e] = slow_entry(cbp, ents[e], col);
                                }
                            }
                            __syncwarp();
                            emitted += listed;
                        }
                        pool_r += T;
                        pool_w += W;
                    }
                }
                }
            next_bundle:
                if (lane == 0) bundle = atomicAdd(&sh->next_bundle, 1);
                bundle = __shfl_sync(FULL, bundle, 0);
                pipe_progress();
            } while (bundle < nb);
        }
        __syncwarp();
        if (lane == 0) pipe_arrive(empty + s);
        s++;
        if (s == n_stages) {
            s = 0;
            par ^= 1u;
        }
    }
    // what is left of the warp's pool stays unused
    for (uint32_t i = pool_r + (uint32_t)lane; i < pool_re; i += WARP) q_index[i] = VQ_INVALID;
#undef GCB_LDS32
}

}  // namespace gcb
