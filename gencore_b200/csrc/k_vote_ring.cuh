// k_vote_ring.cuh — the vote: Pair::computeScore (pair.cpp:88-172) fused with Group::makeConsensus (group.cpp:320-579),
// ONE persistent CTA per SM over a ring of staged tiles (vote_tile.cuh).
//
//   warp 0            producer.  The whole warp fetches the headers of this CTA's next 32 tiles (blockIdx.x, + gridDim.x, ...)
//                     at once; lane 0 then allocates what the tile needs from a ring-buffer arena (a tile takes what it
//                     needs, space is freed in tile order as the oldest tile's `empty` mbarrier completes), writes the slot's
//                     header and issues three bulk asynchronous copies (cp.async.bulk -> UBLKCP) onto the slot's `full`
//                     mbarrier: compact family-side list, VoteRead table, payload slab.  Up to VR_MAX_STAGES tiles in flight.
//   warps 1..15       voters, in G groups (three while six or more of the batch's largest tile fit the arena, else one).  The
//                     CTA's k-th tile belongs to group k % G; every warp of the group visits the group's tiles in order:
//                     waits for `full`, takes bundles of family sides from the tile's counter until none is left, arrives
//                     on `empty`.  A tile of a 16 KB window has about three bundles for the five warps of its group; the
//                     groups decouple the tiles in flight from each other and a warp only walks a third of the tiles.
//   a bundle          32 / L family sides, L lanes each, sixteen columns per lane.  FAST columns (every voter shows the
//                     template's base, no read disagrees with its mate inside the pair overlap, best quality >=
//                     moderateQuality: exactly group.cpp:421-427 under `implied`) are finished in the word: per-column maxima
//                     in 16-bit lanes (VIMNMX3.U16x2 over two reads per iteration), disagreement as OR-accumulated XOR
//                     residues of the raw words; uniform families (every fixed-length library) run a branch-free loop.
//   slow columns      the bundle's slow columns are written into a global queue right away, eight lanes per column (one lane per
//                     read): a 32-byte self-contained header, then per read of the family side its quality, base, mate
//                     quality, mate base and overlap state (4 bytes).  Queue space comes from a per-warp pool reserved with one 64-bit atomic
//                     (records and words in one counter) per ~10 bundles; slow_columns_kernel (k_slow_columns.cuh) decides
//                     the queued columns at full occupancy.  A full queue hands the tile to the generic kernel.  The ring
//                     never waits for a slow column.  (Measured alternatives, profiles/r03_notes.md.)
//   deep tiles        (24 pairs or more per family side on average: few bundles, hundreds of slow columns per tile) keep
//                     their list in the stage; the warp that finishes the tile's last bundle closes it (prefix sums of the
//                     entries' column counts), and ALL voter warps of the tile then decide the columns, 32 at a time, one
//                     thread per column, straight from the staged slab — a deep tile has nothing else for them to do, and
//                     its reads never cross HBM a second time.
#pragma once

#include "k_slow_columns.cuh"

namespace gcb {

#ifndef GCB_VR_THREADS
#define GCB_VR_THREADS 512
#endif
#ifndef GCB_VR_UNROLL
#define GCB_VR_UNROLL 1  // read-loop unrolling (0 = the compiler's choice, which unrolls seven times: 0.198 ms against 0.180 for the
                         // rolled loop on the BASELINE shape — the kernel waits for instructions, not for data; profiles/r03_notes.md)
#endif
constexpr int VR_THREADS = GCB_VR_THREADS;  // warp 0 produces, fifteen warps vote (128 registers per thread)
constexpr int VR_WARPS = VR_THREADS / WARP;
constexpr int VR_VOTERS = VR_WARPS - 1;
constexpr int VR_MAX_STAGES = 16;  // tiles in flight (barrier pairs and stage headers); their bytes come from one ring-buffer arena
constexpr int VR_GROUPS = 3;       // groups of voter warps when the tiles are small
constexpr int VR_GUARD = 4608;     // never allocated, after the arena: the branch-free read loop may read a VoteRead table or a
                                   // slab up to 257 entries / 64 bytes past its end (values unused)

struct __align__(16) RingStage {  // shared memory; the first part is written by the producer before `full` completes
    int64_t out_base0;
    int32_t nfs;          // live family sides; < 0: no more tiles
    int32_t lanes, n_bundles, common_l;
    int32_t p0, tile;
    int32_t ft_off, vr_off, slab_off;  // where the tile's family-side list, VoteRead table and payload slab lie (shared-memory offsets)
    int32_t sl_off;                    // slow-column list: uint32 entries[sl_cap], then their inclusive column counts, then their queue words
    int32_t sl_cap;
    int32_t deep;          // the tile's slow columns are decided here, by all the voter warps together (see the header)
    // the voters' part
    int32_t next_bundle;   // atomic: next bundle to hand out
    int32_t done;          // atomic: bundles finished
    int32_t n_entries;     // atomic: slow-column list entries
    int32_t closed;        // deep tiles: the list is complete and its prefix sums are written
    int32_t drain_total;   // deep tiles: slow columns of the tile
    int32_t next_col;      // deep tiles, atomic: next slow column to decide
    int32_t pad[2];
};
static_assert(sizeof(RingStage) == 96, "stage header size");

// shared-memory map: [barriers][stage headers][producer's header cache][arena]
constexpr int VR_OFF_FULL = 0;                                   // uint64[VR_MAX_STAGES]
constexpr int VR_OFF_EMPTY = 8 * VR_MAX_STAGES;                  // uint64[VR_MAX_STAGES]
constexpr int VR_OFF_HDR = 16 * VR_MAX_STAGES;                   // RingStage[VR_MAX_STAGES]
constexpr int VR_OFF_HCACHE = VR_OFF_HDR + 96 * VR_MAX_STAGES;   // TileHdr2[32]: the producer's next tiles
constexpr int VR_ITEMS = 64;       // slow columns of one bundle that are emitted per round
#ifndef GCB_VR_DEEP_CHUNK
#define GCB_VR_DEEP_CHUNK 32
#endif
constexpr int VR_DEEP_CHUNK = GCB_VR_DEEP_CHUNK;  // slow columns of a deep tile a warp takes at a time (a tile has a few hundred for fifteen warps)
#ifndef GCB_VR_GROUP
#define GCB_VR_GROUP 4
#endif
constexpr int VR_GROUP = GCB_VR_GROUP;  // lanes per slow column in the emission (4: 0.175 ms, 8: 0.180, 2: 0.183, 16: 0.194 on the BASELINE shape)
constexpr int VR_OFF_ITEMS = VR_OFF_HCACHE + 48 * WARP;          // per warp: uint16 codes[VR_ITEMS] (family side in the bundle << 9 | column)
constexpr int VR_OFF_ARENA = (VR_OFF_ITEMS + 2 * VR_ITEMS * VR_WARPS + 127) & ~127;
// a tile's allocation: [FsTile list][VoteRead table][slab + slack][slow-column list + two prefix arrays], each part rounded to 128 bytes
static_assert(VR_OFF_HDR % 16 == 0 && VR_OFF_HCACHE % 16 == 0 && VR_OFF_ARENA % 128 == 0, "ring layout");
GCB_HD uint32_t ring_round128(uint32_t v) { return (v + 127u) & ~127u; }

struct RingCtx {  // what deciding a slow column inside the CTA needs besides the stage
    const BatchView *b;
    const ResultView *r;
    const GenomeView *gv;
    const gcb_options *o;
    RollbackList rb;
};

// One slow column of family side f of a staged (deep) tile, decided from shared memory.  (The pointers are derived from the
// shared-memory symbol inside the function, so that the loads are LDS and not generic loads.)
__device__ __noinline__ void ring_slow_column(const RingCtx &x, int ft_off, int vr_off, int slab_off, int64_t out_base0, int f, int col) {
    GCB_DYN_SMEM(smem);
    const FsTile ft = ((const FsTile *)(smem + ft_off))[f];
    SlowSide fs;
    fs.m = ft.m; fs.l_out = ft.l_out; fs.len = ft.len; fs.tmpl_k = ft.tmpl_k; fs.side = fs_side(ft); fs.flags = ft.flags; fs.slot = ft.slot;
    fs.ref_nib0 = ft.ref_nib0;
    decide_column(*x.b, *x.r, *x.gv, *x.o, x.rb, fs, smem + slab_off + 4 * (int)ft.cbase4, (const VoteRead *)(smem + vr_off) + ft.ent0,
                  x.r->out_payload + out_base0 + 4 * (int64_t)ft.out4, col);
}

__global__ void __launch_bounds__(VR_THREADS, 1) vote_ring_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o, int32_t implied,
                                                                   const TileHdr2 *hdr, const FsTile *fs_tiles, SlowQueue sq, RollbackList rb,
                                                                   int32_t n_tiles, int32_t arena_bytes, const int32_t *max_need) {
    GCB_GRID_DEP();
    GCB_DYN_SMEM(smem);
    if (batch_is_malformed(ws.error_flag)) return;  // (every thread of the grid sees the same flag: the kernels that raise it have finished)
    uint64_t *full = (uint64_t *)(smem + VR_OFF_FULL);
    uint64_t *empty = (uint64_t *)(smem + VR_OFF_EMPTY);
    RingStage *shdr = (RingStage *)(smem + VR_OFF_HDR);
#define GCB_LDS32(off) (*(const uint32_t *)(smem + (off)))
    const int tid = (int)threadIdx.x, lane = lane_id(), warp = tid >> 5;
    // the batch's largest tile (tile_prep2_kernel measured it) decides how the voters are organised: small tiles -> many in
    // flight -> three groups of five warps; tiles that fill the arena -> all fifteen warps on every tile
    const int32_t largest = (int32_t)ring_round128((uint32_t)max(*max_need, 128));
    const int n_groups = (arena_bytes >= 6 * largest && VR_VOTERS % VR_GROUPS == 0) ? VR_GROUPS : 1, wpg = VR_VOTERS / n_groups;
    const int n_stages = VR_MAX_STAGES;
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) {
            pipe_init(full + s, 1);
            pipe_init(empty + s, wpg);  // every voter warp of the tile's group arrives once when it leaves the tile
        }
        pipe_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ---- producer: lane 0 fills the stages, up to n_stages tiles ahead of the voters; the whole warp fetches the
        // headers of this CTA's next 32 tiles at once, so that no tile waits for a header on its way from global memory
        TileHdr2 *hcache = (TileHdr2 *)(smem + VR_OFF_HCACHE);
        int k = 0;
        // ring-buffer arena: the tiles in flight are k_tail .. k-1, their allocations lie between off_of[k_tail] and head (wrapping);
        // slot k % n_stages (barriers, header) once the tile that used it before is released
        int k_tail = 0;
        uint32_t head = 0u, off_of[VR_MAX_STAGES];
        auto release_oldest = [&]() {
            const int s = k_tail % n_stages, use = k_tail / n_stages;
            GCB_TRACE(200 + s);
            while (!pipe_try_wait(empty + s, (uint32_t)(use & 1), 1000u)) pipe_relax(200u);
            GCB_TRACE(300 + s);
            k_tail++;
        };
        auto allocate = [&](uint32_t need) -> uint32_t {
            uint32_t at;
            for (;;) {
                if (k - k_tail == n_stages) {
                    release_oldest();
                    continue;
                }
                if (k == k_tail) {  // nothing in flight
                    at = 0u;
                    break;
                }
                const uint32_t tail = off_of[k_tail % n_stages];
                if (head > tail) {  // free: [head, arena) and [0, tail)
                    if (head + need <= (uint32_t)arena_bytes) { at = head; break; }
                    if (need <= tail) { at = 0u; GCB_COUNT(7, 1); break; }
                } else if (head + need <= tail) {  // free: [head, tail) (head == tail: full)
                    at = head;
                    break;
                }
                release_oldest();
            }
            off_of[k % n_stages] = at;
            head = at + need;
            return at;
        };
        auto fill = [&](RingStage &sh, const TileHdr2 &cur, int32_t tile, uint32_t at) {
            const uint32_t vr_bytes = 32u * (uint32_t)cur.np, ft_bytes = 32u * (uint32_t)max(cur.nfs, 0);
            sh.out_base0 = cur.out_base0;
            sh.nfs = cur.nfs;
            sh.lanes = cur.lanes; sh.n_bundles = cur.n_bundles; sh.common_l = cur.common_l;
            sh.p0 = cur.p0; sh.tile = tile;
            sh.ft_off = VR_OFF_ARENA + (int32_t)at;
            sh.vr_off = sh.ft_off + (int32_t)ring_round128(ft_bytes);
            sh.slab_off = sh.vr_off + (int32_t)ring_round128(vr_bytes);
            sh.sl_off = sh.slab_off + (int32_t)ring_round128((uint32_t)cur.slab_bytes + VT_SLAB_SLACK);
            sh.sl_cap = max(cur.nfs, 0) * cur.lanes;
            sh.deep = tile_is_deep(cur.nfs, cur.np) ? 1 : 0;  // few bundles, long lists of slow columns
            sh.next_bundle = 0;
            sh.done = 0;
            sh.n_entries = 0;
            sh.closed = 0;
            sh.drain_total = 0;
            sh.next_col = 0;
            sh.pad[0] = sh.pad[1] = 0;
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
                    const uint32_t at = allocate((uint32_t)tile_smem_need(cur.nfs, cur.np, cur.slab_bytes, cur.lanes));
                    const int s = k % n_stages;
                    RingStage sh;
                    fill(sh, cur, (int32_t)t, at);
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
        if (lane == 0) {  // one end marker per group: the phase completes with this arrival alone
            for (int g = 0; g < n_groups; g++) {
                const uint32_t at = allocate(128u);
                const int s = k % n_stages;
                TileHdr2 cur;
                cur.out_base0 = 0; cur.slab0 = 0; cur.slab_bytes = 0; cur.p0 = 0; cur.np = 0; cur.nfs = -1; cur.lanes = 1; cur.common_l = 0;
                cur.per_bundle = 32; cur.n_bundles = 0;
                RingStage sh;
                fill(sh, cur, 0, at);
                shdr[s] = sh;
                pipe_expect(full + s, 0u);
                pipe_commit(full + s);
                k++;
            }
        }
        return;
    }

    RingCtx x;
    x.b = &b; x.r = &r; x.gv = &gv; x.o = &o; x.rb = rb;
    // ---- voters
    const uint32_t mod4 = 0x01010101u * (uint32_t)(o.moderate_quality & 0xFF);
    uint32_t pool_r = 0u, pool_re = 0u, pool_w = 0u, pool_we = 0u;  // this warp's reserved records / words of the slow-column queue
    uint16_t *s_item = (uint16_t *)(smem + VR_OFF_ITEMS + 2 * VR_ITEMS * warp);
    const uint32_t sbase = smem_base(smem);
    // lane geometry and masks are kept across tiles while the tile shape (lanes per family side, usual record length) stays
    int cur_L = 0, cur_l = -1, S = 32, sub = 0, j = 0, col0 = 0;
    ChunkMasks cm_common = make_masks(0, 0, 0);
    // this warp's group and the group's tiles: k = group, group + G, ...; tile k lives in slot k % n_stages
    const int group = (warp - 1) % n_groups;
    for (int k = group;; k += n_groups) {
        const int s = k % n_stages;
        const uint32_t par = (uint32_t)((k / n_stages) & 1);
        GCB_TRACE(100 + s);
        pipe_wait(full + s, par, 1000u);
        RingStage *sh = shdr + s;
        const int nfs = sh->nfs;
        if (nfs < 0) break;
        GCB_TRACE(400 + s);
        const int nb = sh->n_bundles;
        const bool deep = sh->deep != 0;
        int bundle = nb;
        if (lane == 0 && *(volatile int32_t *)&sh->next_bundle < nb) bundle = atomicAdd(&sh->next_bundle, 1);  // (no atomic on a drained tile)
        bundle = __shfl_sync(FULL, bundle, 0);
        bool closer = false;
        if (bundle < nb) {
            if (sh->lanes != cur_L) {  // a lane owns sixteen columns; a family side takes L lanes, a bundle 32 / L family sides
                cur_L = sh->lanes;
                S = (int)((32u * ((65535u / (unsigned)cur_L) + 1u)) >> 16);
                sub = (int)(((unsigned)lane * ((65535u / (unsigned)cur_L) + 1u)) >> 16);  // lane / L
                j = lane - sub * cur_L;                                                      // lane % L
                col0 = VT_CHUNK * j;
                cur_l = -1;
            }
            if (sh->common_l != cur_l) {  // the masks of the tile's usual record length
                cur_l = sh->common_l;
                cm_common = make_masks(cur_l, cur_l, col0);
            }
            const int off_slab = sh->slab_off, off_vr = sh->vr_off;
            const FsTile *s_ft = (const FsTile *)(smem + sh->ft_off);
            const VoteRead *s_vr = (const VoteRead *)(smem + off_vr);
            uint8_t *out0 = r.out_payload + sh->out_base0;
            uint32_t *s_list = (uint32_t *)(smem + sh->sl_off);
            const int sb0 = 8 * j;  // the lane's first byte of a record's packed bases
            do {
                const int f = bundle * S + sub;
                const bool have = sub < S && f < nfs;
                FsTile ft = s_ft[have ? f : 0];
                if (!have) ft.mode = SIDE_NONE;
                const int l_out = ft.l_out, len = ft.len;
                const int qbytes = GCB_ALIGN4(l_out), sbytes = GCB_ALIGN4((l_out + 1) >> 1);
                const bool mine = ft.mode != SIDE_NONE && col0 < max(qbytes, 2 * sbytes);  // this lane owns words of the record
                const int m = mine && ft.mode != SIDE_COPY ? (int)ft.m : 0;
                const int mmax = __reduce_max_sync(FULL, m);
                const int cb = off_slab + 4 * (int)ft.cbase4;  // byte offsets into the CTA's shared memory
                const int ento = off_vr + 16 * (int)ft.ent0;
                VoteRead tv = {0, 0, 0, 0, 0, 0, 0, 0};
                uint32_t tbe0 = 0u, tbe1 = 0u;
                int trec = cb;
                if (mine) {
                    tv = s_vr[ft.ent0 + ft.tmpl_k];
                    trec = cb + 4 * (int)tv.own_off4;
                    if (sb0 < sbytes) tbe0 = bswap32(GCB_LDS32(trec + qbytes + sb0));
                    if (sb0 + 4 < sbytes) tbe1 = bswap32(GCB_LDS32(trec + qbytes + sb0 + 4));
                }
                ChunkMasks cm = cm_common;
                if (l_out != cur_l || len != l_out) cm = make_masks(l_out, len, col0);
                if (mine && j == 0 && ft.mode != SIDE_COPY) GCB_COUNT((ft.flags & FS_UNIFORM) ? 4 : 5, 1);
                // per-column maxima live in 16-bit lanes (VIMNMX3.U16x2 is native, a per-byte maximum is seven instructions):
                // mo[k] tracks bytes 1 and 3 of quality word k in the high byte of each half, me[k] bytes 0 and 2 (word << 8)
                uint32_t mo[4] = {0u, 0u, 0u, 0u}, me[4] = {0u, 0u, 0u, 0u}, dis0 = 0u, dis1 = 0u;
                if (ft.flags & FS_UNIFORM) {
                    // hoisted geometry: every voter is read at the template's columns and meets its mate at the same offset,
                    // and every mate's record lies at the same distance from its read's record
                    const int xw = (int)tv.ov_own - col0;   // first column of the lane inside the overlap window
                    const int yw = xw - (int)tv.ov_mate;    // first column of the lane whose mate index is >= 0
                    const int oa = max(max(0, xw), yw), oz = min(min(cm.nvote, xw + (int)tv.ov_len), yw + (int)tv.mate_l);
                    const bool has_ov = tv.ov_len > 0 && oz > oa;
                    const uint32_t om0 = has_ov ? nib_range(oa, oz) : 0u, om1 = has_ov ? nib_range(oa - 8, oz - 8) : 0u;
                    const int ms = 0 - yw, mw0 = ms >> 3;  // the lane's first mate column: word mw0 of the mate's bases, nibble ms & 7
                    const unsigned msh = (unsigned)(ms & 7) * 4u;
                    const uint32_t qbase = sbase + (uint32_t)(cb + col0), sdelta = (uint32_t)(qbytes - col0 + sb0);
                    // mate words relative to the read's own quality chunk.  Words that hold no overlapped column read whatever
                    // lies there (the CTA's shared memory: the lane's window starts at most two words before the mate's bases
                    // and ends inside the slab's slack) and are masked by om0 / om1; lanes without overlap re-read their own chunk.
                    const uint32_t mdelta = has_ov ? (uint32_t)(4 * ((int)tv.mate_off4 - (int)tv.own_off4) + GCB_ALIGN4(tv.mate_l) + 4 * mw0 - col0) : 0u;
                    const uint32_t xt = tv.own_off4;
                    const uint32_t qt = qbase + (xt << 2);
                    const uint32_t t0 = lds32<0>(qt + sdelta), t1 = lds32<4>(qt + sdelta);                              // the template's bases, raw
                    const uint32_t a0 = lds32<0>(qt + mdelta), c0 = lds32<4>(qt + mdelta), e0 = lds32<8>(qt + mdelta);  // its mate's
                    uint32_t d0 = 0u, d1 = 0u, da = 0u, dc = 0u, de = 0u;
                    uint32_t ea = sbase + (uint32_t)ento;
#if GCB_VR_UNROLL == 1
#pragma unroll 1
#elif GCB_VR_UNROLL == 2
#pragma unroll 2
#endif
                    for (int e = 0; e < mmax; e += 2, ea += 32) {
                        uint32_t xa = lds16<0>(ea), xb = lds16<16>(ea);
                        xa = (e < m && xa != VR_NO_VOTE) ? xa : xt;  // reads that do not vote are replaced by the template's own
                        xb = (e + 1 < m && xb != VR_NO_VOTE) ? xb : xt;  // record (maximum and OR are idempotent)
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
                        dis0 |= (__funnelshift_l(C, A, msh) | (bswap32(t0) ^ __funnelshift_l(tc, ta, msh))) & om0;
                        dis1 |= (__funnelshift_l(E, C, msh) | (bswap32(t1) ^ __funnelshift_l(te, tc, msh))) & om1;
                    }
                } else if (m > 0) {
                    for (int e = 0; e < m; e++) {
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
                        if (v.ov_len > 0) {
                            // chunk column k pairs own index rp0+k with mate index k - y.  (Written with subtractions only:
                            // ptxas 12.9 dropped the negation when it folded max(a, max(x, -t)) into one VIMNMX3 on sm_100a —
                            // the PTX was right, the SASS and the B200 were not; profiles/r01_notes.md has the listing.)
                            const int xv = (int)v.ov_own - rp0;  // first chunk column inside the overlap window
                            const int yv = xv - (int)v.ov_mate;  // first chunk column whose mate index is >= 0
                            const int oa = max(max(a, xv), yv);
                            const int oz = min(min(z, xv + (int)v.ov_len), yv + (int)v.mate_l);
                            if (oz > oa) {
                                const uint8_t *mrec = smem + cb + 4 * (int)v.mate_off4;
                                uint32_t mb0, mb1;
                                fetch16b(mrec + GCB_ALIGN4(v.mate_l), GCB_ALIGN4((v.mate_l + 1) >> 1), 0 - yv, mb0, mb1);
                                dis0 |= (be0 ^ mb0) & nib_range(oa, oz);
                                dis1 |= (be1 ^ mb1) & nib_range(oa - 8, oz - 8);
                            }
                        }
                    }
                }
                // ---- what the record gets: qualities = the maxima (fast columns), bases = the template's
                uint32_t slow0 = 0u, slow1 = 0u;  // one bit per slow column (the low bit of its big-endian nibble)
                if (mine) {
                    uint8_t *out = out0 + 4 * (int64_t)ft.out4;
                    uint32_t oq[4];
                    if (ft.mode == SIDE_COPY) {  // group.cpp:73-77: the record itself
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) oq[kk] = col0 + 4 * kk < qbytes ? GCB_LDS32(trec + col0 + 4 * kk) : 0u;
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) oq[kk] = prmt(mo[kk], me[kk], 0x3715u) & cm.vb[kk];  // (the hoisted loop read whole words)
                        GCB_COUNT(2, cm.nvote);
                        if (implied && len == l_out) {
                            const uint32_t lowq0 = nibs_of_flags(bytes_ge_flags(oq[0], mod4) ^ 0x80808080u, bytes_ge_flags(oq[1], mod4) ^ 0x80808080u);
                            const uint32_t lowq1 = nibs_of_flags(bytes_ge_flags(oq[2], mod4) ^ 0x80808080u, bytes_ge_flags(oq[3], mod4) ^ 0x80808080u);
                            slow0 = (dis0 | lowq0) & cm.vn0;
                            slow1 = (dis1 | lowq1) & cm.vn1;
                            slow0 |= slow0 >> 1; slow0 |= slow0 >> 2;  // any differing bit marks the column
                            slow1 |= slow1 >> 1; slow1 |= slow1 >> 2;
                        } else {  // without `implied`, or with columns that are not voted, every column of the record is slow
                            slow0 = nibs_of_bytes(cm.rb[0], cm.rb[1]);
                            slow1 = nibs_of_bytes(cm.rb[2], cm.rb[3]);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
                        if (col0 + 4 * kk < qbytes) *(uint32_t *)(out + col0 + 4 * kk) = oq[kk] & cm.rb[kk];
                    if (sb0 < sbytes) *(uint32_t *)(out + qbytes + sb0) = bswap32(tbe0 & cm.kn0);
                    if (sb0 + 4 < sbytes) *(uint32_t *)(out + qbytes + sb0 + 4) = bswap32(tbe1 & cm.kn1);
                }
                // ---- slow columns: one list entry per lane that found any
                const uint32_t mask16 = nib_flags_to_byte(slow0) | (nib_flags_to_byte(slow1) << 8);
                const unsigned bal = __ballot_sync(FULL, mask16 != 0u);
                if (deep) {
                    // the stage's own list (family side << 21 | lane of the family side << 16 | column mask) ...
                    if (bal != 0u) {
                        int at = 0;
                        if (lane == 0) at = atomicAdd(&sh->n_entries, __popc(bal));
                        at = __shfl_sync(FULL, at, 0);
                        if (mask16 != 0u) s_list[at + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)f << 21) | ((uint32_t)j << 16) | mask16;
                    }
                    // ... and the warp that finishes the tile's last bundle closes the tile
                    __syncwarp();
                    int fin = 0;
                    if (lane == 0) {
                        __threadfence_block();
                        fin = atomicAdd(&sh->done, 1) + 1;
                    }
                    fin = __shfl_sync(FULL, fin, 0);
                    closer = fin == nb;
                } else if (bal != 0u) {
                    // ... or a record per slow column in the global queue: records of one size per bundle, from the warp's own
                    // pool of reserved queue space
                    const uint32_t T = (uint32_t)__reduce_add_sync(FULL, __popc(mask16)), rw = slow_rec_words(mmax), W = T * rw;
                    if (pool_r + T > pool_re || pool_w + W > pool_we) {
                        // a new pool (one 64-bit atomic: records << 32 | words); what is left of the old one stays unused
                        for (uint32_t i = pool_r + (uint32_t)lane; i < pool_re; i += WARP) sq.index[i] = VQ_INVALID;
                        const uint32_t need_r = T > VQ_POOL_RECS ? T : VQ_POOL_RECS, need_w = W > VQ_POOL_WORDS ? W : VQ_POOL_WORDS;
                        unsigned long long base64 = 0ull;
                        if (lane == 0) base64 = atomicAdd(sq.count, ((unsigned long long)need_r << 32) | need_w);
                        base64 = __shfl_sync(FULL, base64, 0);
                        const uint32_t r0 = (uint32_t)(base64 >> 32), w0 = (uint32_t)base64;
                        if ((unsigned long long)r0 + need_r <= sq.cap_recs && (unsigned long long)w0 + need_w <= sq.cap_words) {
                            pool_r = r0; pool_re = r0 + need_r;
                            pool_w = w0; pool_we = w0 + need_w;
                        } else {  // the queue is full: the reserved index entries are marked unused, the pool stays empty
                            for (uint32_t i = r0 + (uint32_t)lane; i < r0 + need_r && i < sq.cap_recs; i += WARP) sq.index[i] = VQ_INVALID;
                            pool_r = pool_re = pool_w = pool_we = 0u;
                        }
                    }
                    if (pool_r + T > pool_re) {
                        // no queue space: the generic kernel redoes the whole tile from the payload (it runs after
                        // slow_columns_kernel and vote_rollback_kernel)
                        if (lane == 0 && atomicExch(&sh->closed, 1) == 0) {
                            ws.generic_tiles[atomicAdd(ws.generic_count, 1)] = ~sh->tile;
                            GCB_COUNT(1, 1);
                        }
                    } else {
                        // rounds of at most VR_ITEMS columns: every lane lists one of its columns per ballot, then eight lanes per
                        // column write the column's entries, one lane per read (a lane that writes its own columns alone keeps
                        // the other lanes waiting: slow columns are rare, the warp has two or three lanes with any)
                        uint32_t mk = mask16, emitted = 0u;
                        while (emitted < T) {
                            uint32_t listed = 0u;
                            while (listed + WARP <= (uint32_t)VR_ITEMS) {
                                const unsigned lb = __ballot_sync(FULL, mk != 0u);
                                if (lb == 0u) break;
                                if (mk != 0u) {
                                    const int bit = __ffs((int)mk) - 1;  // bit 8 * w + i of the mask = column 8 * w + 7 - i of the lane's sixteen
                                    mk &= mk - 1u;
                                    s_item[listed + __popc(lb & ((1u << lane) - 1u))] = (uint16_t)((sub << 9) | (col0 + (bit & 8) + 7 - (bit & 7)));
                                }
                                listed += __popc(lb);
                            }
                            __syncwarp();
                            const int g = lane / VR_GROUP, gl = lane % VR_GROUP;
                            for (uint32_t i = (uint32_t)g; i < listed; i += WARP / VR_GROUP) {
                                const uint32_t code = s_item[i], wofs = pool_w + (emitted + i) * rw;
                                const int col = (int)(code & 511u);
                                const FsTile fti = s_ft[bundle * S + (int)(code >> 9)];
                                const uint8_t *cbp = smem + off_slab + 4 * (int)fti.cbase4;
                                const VoteRead *ents = s_vr + fti.ent0;
                                uint32_t *rec = sq.words + wofs;
                                const int mi = (int)fti.m;
                                if (gl == 0) {
                                    sq.index[pool_r + emitted + i] = wofs;
                                    slow_write_header(rec, fti, col, sh->out_base0 + 4 * (int64_t)fti.out4);
                                }
                                if ((fti.flags & FS_UNIFORM) && col < (int)fti.len) {
                                    // the column's place in pair.cpp:121-170 is the same for every read of a uniform family
                                    const VoteRead tvi = ents[fti.tmpl_k];
                                    const bool info = tvi.ov_len != VR_NO_OVERLAP_INFO;
                                    const int kq = col - (int)tvi.ov_own, mp = (int)tvi.ov_mate + kq;
                                    const bool inwin = info && kq >= 0 && kq < (int)tvi.ov_len;
                                    const bool mvalid = inwin && mp >= 0 && mp < (int)tvi.mate_l;
                                    const uint32_t tag = ((!info ? SE_NO_INFO : !inwin ? SE_PLAIN : mvalid ? SE_MATE : SE_NO_MATE_BASE) << 24) | SE_VOTES;
                                    const int soff = GCB_ALIGN4(fti.l_out) + (col >> 1), nsh = (col & 1) ? 0 : 4;
                                    const int mpi = mvalid ? mp : 0;
                                    const int msoff = GCB_ALIGN4(tvi.mate_l) + (mpi >> 1), mnsh = (mpi & 1) ? 0 : 4;
                                    for (int e = gl; e < mi; e += VR_GROUP) {
                                        const uint32_t w = *(const uint32_t *)(ents + e);  // own_off4 | mate_off4 << 16
                                        uint32_t ent = 0u;
                                        if ((w & 0xFFFFu) != VR_NO_VOTE) {
                                            const uint8_t *p = cbp + 4 * (int)(w & 0xFFFFu);
                                            ent = (uint32_t)p[col] | ((((uint32_t)p[soff] >> nsh) & 0xFu) << 16) | tag;
                                            if (mvalid) {
                                                const uint8_t *q = cbp + 4 * (int)(w >> 16);
                                                ent |= ((uint32_t)q[mpi] << 8) | ((((uint32_t)q[msoff] >> mnsh) & 0xFu) << 20);
                                            }
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
                if (lane == 0) bundle = atomicAdd(&sh->next_bundle, 1);
                bundle = __shfl_sync(FULL, bundle, 0);
                pipe_progress();
            } while (bundle < nb);
        }
        if (closer) {
            // every bundle of the (deep) tile is done: prefix sums of the entries' column counts, then the list is open to every warp
            __threadfence_block();
            const int n = *(volatile int32_t *)&sh->n_entries;
            uint32_t *s_list = (uint32_t *)(smem + sh->sl_off), *s_pf = s_list + sh->sl_cap;
            int run = 0;
            for (int base = 0; base < n; base += WARP) {
                const int i = base + lane;
                int incl = i < n ? __popc(s_list[i] & 0xFFFFu) : 0;
                for (int off = 1; off < WARP; off <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, off);
                    if (lane >= off) incl += v;
                }
                if (i < n) s_pf[i] = (uint32_t)(run + incl);
                run += __shfl_sync(FULL, incl, WARP - 1);
            }
            __syncwarp();
            if (lane == 0) {
                sh->drain_total = run;
                __threadfence_block();
                *(volatile int32_t *)&sh->closed = 1;
            }
        }
        if (deep) {
            // every voter warp waits for the tile to be closed and then decides slow columns, 32 at a time, one thread per
            // column, straight from the staged slab: a deep tile has hundreds of them and nothing else for the warps to do
            while (*(volatile int32_t *)&sh->closed == 0) pipe_relax(400u);  // (thirteen warps poll while two vote; the length of the nap does not matter: 100 to 1000 ns measured)
            __threadfence_block();
            const int total = *(volatile int32_t *)&sh->drain_total, n = *(volatile int32_t *)&sh->n_entries;
            const uint32_t *s_list = (const uint32_t *)(smem + sh->sl_off), *s_pf = s_list + sh->sl_cap;
            const int ft_off = sh->ft_off, vr_off = sh->vr_off, slab_off = sh->slab_off;
            const int64_t out_base0 = sh->out_base0;
            for (;;) {
                int c0 = 0;
                if (lane == 0) c0 = atomicAdd(&sh->next_col, VR_DEEP_CHUNK);
                c0 = __shfl_sync(FULL, c0, 0);
                if (c0 >= total) break;
                const int idx = c0 + lane;
                if (lane < VR_DEEP_CHUNK && idx < total) {
                    int lo = 0, hi = n - 1;
                    while (lo < hi) {  // the first entry whose inclusive column count exceeds idx
                        const int mid = (lo + hi) >> 1;
                        if ((int)s_pf[mid] > idx) hi = mid;
                        else lo = mid + 1;
                    }
                    const uint32_t code = s_list[lo];
                    uint32_t mask = code & 0xFFFFu;
                    const int rank = idx - ((int)s_pf[lo] - __popc(mask));
                    for (int q = 0; q < rank; q++) mask &= mask - 1u;
                    const int bit = __ffs((int)mask) - 1;
                    const int col = VT_CHUNK * (int)((code >> 16) & 31u) + (bit & 8) + 7 - (bit & 7);
                    ring_slow_column(x, ft_off, vr_off, slab_off, out_base0, (int)(code >> 21), col);
                }
                __syncwarp();
                pipe_progress();
            }
        }
        __syncwarp();
        if (lane == 0) pipe_arrive(empty + s);
    }
    for (uint32_t i = pool_r + (uint32_t)lane; i < pool_re; i += WARP) sq.index[i] = VQ_INVALID;  // what is left of the warp's pool
#undef GCB_LDS32
}

}  // namespace gcb
