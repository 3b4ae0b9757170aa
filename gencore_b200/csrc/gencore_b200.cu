// gencore_b200.cu — the C ABI of libgencore_b200.so (include/gencore_b200.h) and the launch sequence
// of one batch.  Host code only sizes buffers, moves bytes and launches; all arithmetic of the
// reference's hot path lives in the kernels:
//   umi_group_kernel -> select_template_kernel -> scan_local/scan_blocks -> tile_prep2_kernel -> vote_ring_kernel
//   -> slow_columns_kernel -> vote_rollback_kernel -> score_vote_kernel (the tiles that do not fit the ring's arena) -> duplex_kernel
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <new>
#include <vector>

#include "k_cluster_stats.cuh"
#include "k_duplex.cuh"
#include "k_group_select.cuh"
#include "k_score_vote.cuh"
#include "k_stat_depth.cuh"
#include "k_fasta_pack.cuh"
#include "k_umi_extract.cuh"
#include "k_vote_ring.cuh"

using namespace gcb;

constexpr int GCB_MAX_CHUNKS = 16;
constexpr int64_t GCB_CHUNK_BYTES = 48ll << 20;  // payload per pipeline chunk of gcb_consensus_batch

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct gcb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    gcb_options opt;
    char err[256] = {0};
    int64_t launches = 0;
    // genome
    GenomeView genome = {nullptr, nullptr, nullptr, 0, 0};
    DevBuf g_packed, g_off, g_len;
    // workspace (grow-only)
    DevBuf w_members, w_group_off, w_scratch, w_rrp, w_flags, w_mode, w_hasumi, w_overlap, w_slab, w_cob, w_coo, w_scan, w_err, w_tiles;
    DevBuf w_vr, w_fs, w_gtiles, w_gcount;
    DevBuf w_fstiles, w_thdr2, w_need;   // the tiles' compact family-side lists, their headers, the largest tile's shared-memory need (per chunk)
    DevBuf w_rb_list, w_rb_count;        // rollback candidates (per chunk: one counter)
    DevBuf w_stats;                      // gcb_cluster_stats accumulator (cluster_stats_kernel)
    DevBuf w_sq_count, w_sq_words, w_sq_index;  // slow-column queue (per chunk: one counter)
    int64_t slow_queue_bytes = 0;        // 0 = sized from the payload
    uint32_t sq_cap_words = 0, sq_cap_recs = 0;
    int ring_window_shift = 0;           // 0 = chosen by plan_tiles; 14 / 15 = forced (tuning)
    int group_lanes = 0;                 // lanes per cluster in umi_group / select_template (0 = by mean cluster size)
    int force_generic = 0;               // tests: every tile goes to the generic kernel
    int trace_e2e = 0;                   // gcb_set_debug key 6: gcb_consensus_batch prints where its time goes (stderr)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    int n_sms = 148;
    // device mirror of a host batch / result (gcb_consensus_batch)
    DevBuf d_pair_off, d_cref, d_cflags, d_umi, d_reads, d_cigar, d_payload;
    DevBuf d_pair_group, d_ngroups, d_groups, d_out, d_out_bytes;
    DevBuf s_tid, s_pos, s_len, s_off, s_depth;  // gcb_stat_depth
    DevBuf u_names, u_off, u_out, u_status;  // gcb_extract_umi
    DevBuf f_text, f_anchor, f_cnt, f_hpos, f_hbase, f_flag, f_coff, f_out;  // gcb_pack_fasta
    // gcb_consensus_batch pipelines chunks of clusters: copies in, kernels and copies out run on three streams
    cudaStream_t h2d = nullptr, d2h = nullptr, d2h_out = nullptr;  // (d2h_out: the consensus records, whose sizes the host learns chunk by chunk)
    cudaEvent_t ev_in[GCB_MAX_CHUNKS] = {nullptr}, ev_done[GCB_MAX_CHUNKS] = {nullptr}, ev_out[GCB_MAX_CHUNKS] = {nullptr};
    int64_t *h_totals = nullptr;  // pinned: cumulative consensus bytes after every chunk
    int32_t *h_flag = nullptr;    // pinned: the device error flag
    int64_t chunk_bytes = GCB_CHUNK_BYTES;
};

namespace {

int fail(gcb_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        if (e != cudaSuccess) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, cudaGetErrorString(e));
        else snprintf(ctx->err, sizeof ctx->err, "%s", what);
    }
    return code;
}

#define GCB_CUDA(ctx, call)                                              \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return fail((ctx), GCB_ERR_CUDA, #call, e_); \
    } while (0)

int reserve(gcb_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return GCB_OK;
    if (b.p) GCB_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    GCB_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return GCB_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// Tile geometry of the vote for one batch: the payload window whose clusters form a tile, and the shared-memory arena of
// the ring kernel, which must hold the largest such tile (window + the largest cluster) at least once.
struct TilePlan {
    int32_t window, window_shift, slab_cap;
    int32_t ring;         // vote_ring_kernel takes the tiles (0: clusters too large for the arena, the generic kernel takes all)
    int32_t arena, smem;  // bytes
};
TilePlan plan_tiles(int32_t max_cluster_bytes, int ring_window_shift = 0) {
    const int32_t KB = 1024, budget = 227 * KB;
    const int32_t maxc = max_cluster_bytes > 0 ? ((max_cluster_bytes + 127) & ~127) : 16 * KB;
    TilePlan p;
    memset(&p, 0, sizeof p);
    p.arena = (budget - 1 * KB - VR_OFF_ARENA - VR_GUARD) & ~127;
    p.smem = VR_OFF_ARENA + p.arena + VR_GUARD;
    // 32 KB windows: a tile is about a window plus half a cluster; five or six of them are in flight in the ring kernel's arena and
    // all fifteen voter warps walk every tile.  A batch whose largest cluster nearly fills the arena still gets one tile in flight.
    // (16 KB windows, with the voters in three groups, measured slower on the BASELINE shapes: twice the tiles for the one
    // producer lane and for every warp's walk; gcb_set_debug key 2 selects them.)
    {
        const int shift = ring_window_shift ? ring_window_shift : 15;
        p.window_shift = shift;
        p.window = 1 << shift;
        p.slab_cap = p.window + maxc;
        if (p.slab_cap > VT_MAX_SLAB) p.slab_cap = VT_MAX_SLAB;
        // a tile of the largest slab with the tables of a family of pairs that fills it
        const int32_t tables = 4 * 128 + 12 * KB;
        const int32_t largest = p.window + maxc + VT_SLAB_SLACK + tables;
        if (p.window + maxc <= VT_MAX_SLAB && p.arena >= largest) {
            p.ring = 1;
            return p;
        }
    }
    // clusters too large for a stage: every tile goes to the generic kernel (no size limits)
    p.window_shift = ring_window_shift ? ring_window_shift : 15;
    p.window = 1 << p.window_shift;
    p.slab_cap = 0;
    p.ring = 0;
    return p;
}

// group.cpp:421-427 holds for a column whose reads all agree and whose best quality is >= moderateQuality
// without looking at the scores iff every score such a column can see is positive and the best read's is >= -c
int32_t fast_path_implied(const gcb_options &o) {
    const int sh = (signed char)o.score_high, sm = (signed char)o.score_moderate, sl = (signed char)o.score_low, sb = (signed char)o.score_bad;
    const int mn = sh < sm ? (sh < sl ? (sh < sb ? sh : sb) : (sl < sb ? sl : sb)) : (sm < sl ? (sm < sb ? sm : sb) : (sl < sb ? sl : sb));
    if (mn <= 0 || mn + 4 > 127) return 0;
    if (sh < o.base_score_req || sm < o.base_score_req || mn + 4 < o.base_score_req) return 0;
    if (o.moderate_quality < 0 || o.moderate_quality > 128) return 0;  // (bytes_ge_flags compares bytes against thresholds <= 128)
    return 1;
}

int reserve_workspace(gcb_ctx *ctx, int64_t n_pairs, int64_t n_clusters, int64_t n_tiles, int64_t payload_bytes, Workspace &ws) {
    int rc;
    const int64_t n_scan = (n_clusters + SCAN_BLOCK - 1) / SCAN_BLOCK;
#define GCB_RES(buf, bytes) if ((rc = reserve(ctx, ctx->buf, (size_t)(bytes))) != GCB_OK) return rc
    GCB_RES(w_members, n_pairs * 4);
    GCB_RES(w_group_off, n_pairs * 4);
    GCB_RES(w_scratch, n_pairs * 8);
    GCB_RES(w_rrp, n_pairs * 8);
    GCB_RES(w_flags, n_pairs * 2);
    GCB_RES(w_mode, n_pairs * 2);
    GCB_RES(w_hasumi, n_clusters);
    GCB_RES(w_overlap, n_pairs * sizeof(PairOverlap));
    GCB_RES(w_slab, (n_clusters + GCB_MAX_CHUNKS + 1) * 8);
    GCB_RES(w_cob, n_clusters * 8);
    GCB_RES(w_coo, n_clusters * 8);
    GCB_RES(w_scan, (n_scan + 2 * GCB_MAX_CHUNKS + 1) * 8);
    GCB_RES(w_err, 4);
    GCB_RES(w_tiles, (n_tiles + 2 * GCB_MAX_CHUNKS + 1) * sizeof(TileDir));
    GCB_RES(w_vr, 2 * n_pairs * sizeof(VoteRead));
    GCB_RES(w_fs, 2 * n_pairs * sizeof(FsDesc));
    GCB_RES(w_gtiles, (n_tiles + 2 * GCB_MAX_CHUNKS + 1) * 4);
    GCB_RES(w_gcount, 4 * GCB_MAX_CHUNKS);
    GCB_RES(w_fstiles, 2 * n_pairs * sizeof(FsTile));
    GCB_RES(w_thdr2, (n_tiles + 2 * GCB_MAX_CHUNKS + 1) * sizeof(TileHdr2));
    GCB_RES(w_need, 4 * GCB_MAX_CHUNKS);
    GCB_RES(w_rb_list, 2 * n_pairs * 4);
    GCB_RES(w_rb_count, 4 * GCB_MAX_CHUNKS);
    if (!ctx->w_stats.p) {
        GCB_RES(w_stats, sizeof(gcb_cluster_stats));
        GCB_CUDA(ctx, cudaMemsetAsync(ctx->w_stats.p, 0, sizeof(gcb_cluster_stats), ctx->stream));
        GCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // (the first batch may run on a caller's stream)
    }
    {   // slow-column queue: a clean shallow library queues about 0.08 bytes per payload byte, a noisy one of depth 30 (1 % errors:
        // a third of its columns are slow) about 1.5, a vote whose every column is slow (options outside fast_path_implied) about
        // 4; deep families never use it (their tiles are decided in place); tiles whose columns do not fit are redone by the
        // generic kernel
        int64_t qbytes = ctx->slow_queue_bytes > 0 ? ctx->slow_queue_bytes
                         : fast_path_implied(ctx->opt) ? 2 * payload_bytes + (16ll << 20) : 5 * payload_bytes + (16ll << 20);
        int64_t cap_words = qbytes / 4;
        if (cap_words > 0xFFFFFFF0ll) cap_words = 0xFFFFFFF0ll;
        cap_words &= ~3ll;
        if (cap_words < 64) cap_words = 64;
        const int64_t cap_recs = cap_words / 12;  // the smallest record is 12 words
        GCB_RES(w_sq_count, 8 * GCB_MAX_CHUNKS);
        GCB_RES(w_sq_words, 4 * cap_words);
        GCB_RES(w_sq_index, 4 * cap_recs);
        ctx->sq_cap_words = (uint32_t)cap_words;
        ctx->sq_cap_recs = (uint32_t)cap_recs;
    }
#undef GCB_RES
    ws.members = (int32_t *)ctx->w_members.p;
    ws.group_off = (int32_t *)ctx->w_group_off.p;
    ws.scratch = (int32_t *)ctx->w_scratch.p;
    ws.right_ref_pos = (int32_t *)ctx->w_rrp.p;
    ws.vote_flags = (uint8_t *)ctx->w_flags.p;
    ws.side_mode = (uint8_t *)ctx->w_mode.p;
    ws.cluster_has_umi = (uint8_t *)ctx->w_hasumi.p;
    ws.overlap = (PairOverlap *)ctx->w_overlap.p;
    ws.slab_off = (int64_t *)ctx->w_slab.p;
    ws.cluster_out_bytes = (int64_t *)ctx->w_cob.p;
    ws.cluster_out_off = (int64_t *)ctx->w_coo.p;
    ws.scan_block = (int64_t *)ctx->w_scan.p;
    ws.error_flag = (int32_t *)ctx->w_err.p;
    ws.tile_dir = (TileDir *)ctx->w_tiles.p;
    ws.vote_reads = (VoteRead *)ctx->w_vr.p;
    ws.fs_desc = (FsDesc *)ctx->w_fs.p;
    ws.generic_tiles = (int32_t *)ctx->w_gtiles.p;
    ws.generic_count = (int32_t *)ctx->w_gcount.p;
    return GCB_OK;
}

// A contiguous run of clusters processed as one unit (the whole batch, or one pipeline chunk of it).
struct ViewRange {
    int32_t c0, c1;        // clusters
    int32_t p0, p1;        // pairs
    int64_t s0, s1;        // payload bytes
    int64_t tile_base;     // first entry of this view in tile_dir / generic_tiles
    int64_t scan_base;     // first entry of this view in scan_block
    int32_t index;         // chunk number (generic_count slot)
};

// The kernels of the path over one view.  Every array of `batch` / `result` is a device pointer to the WHOLE
// batch; per-pair arrays are indexed absolutely, per-cluster arrays are rebased to the view here.
int launch_stages(gcb_ctx *ctx, const gcb_batch &batch, const gcb_result &result, const Workspace &ws0, const TilePlan &plan,
                  const ViewRange &v, uint32_t stages, cudaStream_t stream, const int64_t *carry_in, int64_t *total_out) {
    const int32_t nc = v.c1 - v.c0;
    const int64_t n_tiles = (v.s1 - v.s0 + plan.window - 1) / plan.window;
    BatchView b = {nc, v.p1, batch.umi_words, batch.cluster_pair_off + v.c0, batch.cluster_ref + v.c0, batch.cluster_flags + v.c0,
                   batch.umi, batch.reads, batch.cigar, batch.payload, v.s1, v.s0, batch.n_cigar_ops};
    ResultView r = {result.pair_group, result.cluster_n_groups + v.c0, result.groups, result.out_payload, result.out_capacity, total_out};
    Workspace ws = ws0;
    ws.cluster_has_umi += v.c0;
    ws.slab_off += v.c0 + v.index;  // every view writes one entry past its last cluster
    ws.cluster_out_bytes += v.c0;
    ws.cluster_out_off += v.c0;
    ws.scan_block += v.scan_base;
    ws.tile_dir += v.tile_base;
    ws.generic_tiles += v.tile_base;
    ws.generic_count += v.index;
    if (nc == 0) {
        if (stages & GCB_STAGE_SELECT_TEMPLATE) {
            if (carry_in) GCB_CUDA(ctx, cudaMemcpyAsync(total_out, carry_in, 8, cudaMemcpyDeviceToDevice, stream));
            else GCB_CUDA(ctx, cudaMemsetAsync(total_out, 0, 8, stream));
        }
        return GCB_OK;
    }
    // umi_group_kernel / select_template_kernel give GS lanes to a cluster and keep a cluster of at most GS pairs in registers:
    // wide enough for most of the batch's clusters to take that path
    int gs = ctx->group_lanes;
    if (gs != 8 && gs != 16 && gs != 32) {
        const int64_t avg = (int64_t)(v.p1 - v.p0) / nc;
        gs = avg <= 4 ? 8 : avg <= 12 ? 16 : 32;
    }
    const int clusters_per_cta = (GROUP_THREADS / WARP) * (WARP / gs);
    const unsigned grid_clusters = (unsigned)((nc + clusters_per_cta - 1) / clusters_per_cta);
    // select_template_kernel keeps a cluster of at most GS pairs in registers: sixteen lanes unless the clusters are tiny or huge
    int gs_sel = ctx->group_lanes;
    if (gs_sel != 8 && gs_sel != 16 && gs_sel != 32) {
        const int64_t avg = (int64_t)(v.p1 - v.p0) / nc;
        gs_sel = avg <= 4 ? 8 : avg <= 12 ? 16 : 32;
    }
    const int sel_per_cta = (GROUP_THREADS / WARP) * (WARP / gs_sel);
    const unsigned grid_sel = (unsigned)((nc + sel_per_cta - 1) / sel_per_cta);
#define GCB_UMI_LAUNCH(NW)                                                                                                                    \
    do {                                                                                                                                      \
        if (gs == 8) GCB_LAUNCH((umi_group_kernel<NW, 8>), dim3(grid_clusters), dim3(GROUP_THREADS), 0, stream, b, r, ws, plan.window_shift, (int32_t)n_tiles);        \
        else if (gs == 16) GCB_LAUNCH((umi_group_kernel<NW, 16>), dim3(grid_clusters), dim3(GROUP_THREADS), 0, stream, b, r, ws, plan.window_shift, (int32_t)n_tiles); \
        else GCB_LAUNCH((umi_group_kernel<NW, 32>), dim3(grid_clusters), dim3(GROUP_THREADS), 0, stream, b, r, ws, plan.window_shift, (int32_t)n_tiles);               \
    } while (0)
    if (stages & GCB_STAGE_UMI_GROUP) {
        if (batch.umi_words == 1) GCB_UMI_LAUNCH(1);
        else if (batch.umi_words == 2) GCB_UMI_LAUNCH(2);
        else if (batch.umi_words == 3) GCB_UMI_LAUNCH(3);
        else GCB_UMI_LAUNCH(4);
        ctx->launches++;
    }
#undef GCB_UMI_LAUNCH
    if (stages & GCB_STAGE_SELECT_TEMPLATE) {
        const int32_t n_scan = (int32_t)((nc + SCAN_BLOCK - 1) / SCAN_BLOCK);
        if (gs_sel == 8) GCB_LAUNCH(select_template_kernel<8>, dim3(grid_sel), dim3(GROUP_THREADS), 0, stream, b, r, ws, ctx->genome, ctx->opt);
        else if (gs_sel == 16) GCB_LAUNCH(select_template_kernel<16>, dim3(grid_sel), dim3(GROUP_THREADS), 0, stream, b, r, ws, ctx->genome, ctx->opt);
        else GCB_LAUNCH(select_template_kernel<32>, dim3(grid_sel), dim3(GROUP_THREADS), 0, stream, b, r, ws, ctx->genome, ctx->opt);
        GCB_LAUNCH(scan_local_kernel, dim3((unsigned)n_scan), dim3(SCAN_THREADS), 0, stream, ws, nc);
        GCB_LAUNCH(scan_blocks_kernel, dim3(1), dim3(WARP), 0, stream, ws, n_scan, total_out, result.out_capacity, carry_in);
        ctx->launches += 3;
    }
    // measurement: the vote's three parts can be run one at a time (tile preparation; the ring kernel; rollback + generic)
    const bool run_prep = (stages & (GCB_STAGE_SCORE_VOTE | GCB_STAGE_VOTE_PREP_ONLY)) != 0;
    const bool run_vote = (stages & (GCB_STAGE_SCORE_VOTE | GCB_STAGE_VOTE_ONLY)) != 0;
    const bool run_fast = run_vote || (stages & GCB_STAGE_VOTE_FAST_ONLY) != 0;
    const bool run_rest = run_vote || (stages & GCB_STAGE_VOTE_REST_ONLY) != 0;
    if ((run_prep || run_fast || run_rest) && n_tiles > 0) {
        TileHdr2 *thdr = (TileHdr2 *)ctx->w_thdr2.p + v.tile_base;
        FsTile *fst = (FsTile *)ctx->w_fstiles.p;
        int32_t *max_need = (int32_t *)ctx->w_need.p + v.index;
        // chunks of one batch run one after another on the stream; every chunk has its own rollback counter and list range
        RollbackList rb;
        rb.list = (int32_t *)ctx->w_rb_list.p + 2 * (size_t)v.p0;
        rb.count = (int32_t *)ctx->w_rb_count.p + v.index;
        rb.cap = 2 * (v.p1 - v.p0);
        // ... and its own queue counter (the queue itself is shared: the chunks' votes never overlap)
        SlowQueue sq;
        sq.count = (unsigned long long *)ctx->w_sq_count.p + v.index;
        sq.words = (uint32_t *)ctx->w_sq_words.p;
        sq.index = (uint32_t *)ctx->w_sq_index.p;
        sq.cap_words = ctx->sq_cap_words;
        sq.cap_recs = ctx->sq_cap_recs;
        if (run_prep) {
            GCB_CUDA(ctx, cudaMemsetAsync(ws.generic_count, 0, 4, stream));
            GCB_CUDA(ctx, cudaMemsetAsync(max_need, 0, 4, stream));
            GCB_LAUNCH(tile_prep2_kernel, dim3((unsigned)((n_tiles + VS_PREP_THREADS / WARP - 1) / (VS_PREP_THREADS / WARP))), dim3(VS_PREP_THREADS), 0,
                       stream, b, r, ws, plan.slab_cap, plan.arena, thdr, fst, max_need, (int32_t)n_tiles, (int32_t)(ctx->force_generic || !plan.ring));
            ctx->launches++;
        }
        if (run_fast && plan.ring && !ctx->force_generic) {
            GCB_CUDA(ctx, cudaMemsetAsync(rb.count, 0, 4, stream));
            GCB_CUDA(ctx, cudaMemsetAsync(sq.count, 0, 8, stream));
            const unsigned ring_grid = (unsigned)(n_tiles < ctx->n_sms ? n_tiles : ctx->n_sms);
            GCB_LAUNCH(vote_ring_kernel, dim3(ring_grid), dim3(VR_THREADS), plan.smem, stream, b, r, ws, ctx->genome, ctx->opt,
                       fast_path_implied(ctx->opt), (const TileHdr2 *)thdr, (const FsTile *)fst, sq, rb, (int32_t)n_tiles, plan.arena,
                       (const int32_t *)max_need);
            ctx->launches++;
        }
        if (run_rest) {
            if (plan.ring && !ctx->force_generic) {
                GCB_LAUNCH(slow_columns_kernel, dim3(VQ_SLOW_CTAS), dim3(VQ_SLOW_THREADS), 0, stream, b, r, ws, ctx->genome, ctx->opt, sq, rb);
                ctx->launches++;
                GCB_LAUNCH(vote_rollback_kernel, dim3(VQ_FINAL_CTAS), dim3(VQ_FINAL_THREADS), 0, stream, b, r, ws, ctx->opt, rb, v.p0, v.p1);
                ctx->launches++;
            }
            // the tiles tile_prep2_kernel handed over (an empty list costs a few microseconds)
            const unsigned generic_grid = (unsigned)(n_tiles < 2 * 148 ? n_tiles : 2 * 148);
            GCB_LAUNCH(score_vote_kernel, dim3(generic_grid), dim3(VOTE_THREADS), VOTE_SMEM, stream, b, r, ws, ctx->genome, ctx->opt);
            ctx->launches++;
        }
    }
    if (stages & GCB_STAGE_DUPLEX) {
        // lanes per cluster = families it pairs up in registers (the tuning knob of the other two kernels applies here too)
        const int gs_dup = ctx->group_lanes == 8 ? 8 : (ctx->group_lanes == 16 || ctx->group_lanes == 32) ? 16 : (int64_t)(v.p1 - v.p0) / nc <= 12 ? 8 : 16;
        const dim3 grid_dup((unsigned)((gs_dup * (int64_t)nc + DUPLEX_THREADS - 1) / DUPLEX_THREADS));
        if (gs_dup == 8) GCB_LAUNCH(duplex_kernel<8>, grid_dup, dim3(DUPLEX_THREADS), 0, stream, b, r, ws, ctx->opt);
        else GCB_LAUNCH(duplex_kernel<16>, grid_dup, dim3(DUPLEX_THREADS), 0, stream, b, r, ws, ctx->opt);
        ctx->launches++;
        // ... and the Stats side effects of the clusters' verdicts
        const unsigned stats_grid = (unsigned)((nc + STATS_THREADS - 1) / STATS_THREADS);
        GCB_LAUNCH(cluster_stats_kernel, dim3(stats_grid < 4u * 148u ? stats_grid : 4u * 148u), dim3(STATS_THREADS), 0, stream, b, r, ws,
                   (unsigned long long *)ctx->w_stats.p);
        ctx->launches++;
    }
    GCB_CUDA(ctx, cudaGetLastError());
    return GCB_OK;
}

}  // namespace

extern "C" {

int gcb_abi_version(void) { return GCB_ABI_VERSION; }

void gcb_default_options(gcb_options *o) {  // Options::Options, options.cpp:4-40
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->duplex_mismatch_threshold = 2;
    o->cluster_size_req = 1;
    o->base_score_req = 6;
    o->high_quality = 30;
    o->moderate_quality = 20;
    o->low_quality = 15;
    o->score_high = 8;
    o->score_moderate = 6;
    o->score_low = 4;
    o->score_bad = 2;
    o->skip_low_complexity_cluster_threshold = 1000;
    o->score_percent_req = 0.8;
}

int gcb_create(const gcb_options *opt, int device, gcb_ctx **out) {
    if (!out) return GCB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return GCB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GCB_ERR_NO_DEVICE;
    if (prop.major != 10) return GCB_ERR_NO_DEVICE;  // the kernels exist for sm_100a only; there is no other path
    gcb_ctx *ctx = new (std::nothrow) gcb_ctx();
    if (!ctx) return GCB_ERR_ARG;
    ctx->device = device;
    ctx->n_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if (opt) ctx->opt = *opt;
    else gcb_default_options(&ctx->opt);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return GCB_ERR_CUDA;
    }
    bool ok = cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->d2h_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->h_totals, 8 * (GCB_MAX_CHUNKS + 1)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->h_flag, 8) == cudaSuccess;
    for (int k = 0; ok && k < GCB_MAX_CHUNKS; k++)
        ok = cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_out[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        gcb_destroy(ctx);
        return GCB_ERR_CUDA;
    }
    if (cudaFuncSetAttribute(score_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VOTE_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(vote_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        gcb_destroy(ctx);
        return GCB_ERR_CUDA;
    }
    *out = ctx;
    return GCB_OK;
}

void gcb_destroy(gcb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->h2d) cudaStreamSynchronize(ctx->h2d);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->d2h) cudaStreamSynchronize(ctx->d2h);
    if (ctx->d2h_out) cudaStreamSynchronize(ctx->d2h_out);
    DevBuf *all[] = {&ctx->g_packed, &ctx->g_off, &ctx->g_len, &ctx->w_members, &ctx->w_group_off, &ctx->w_scratch, &ctx->w_rrp,
                     &ctx->w_flags, &ctx->w_mode, &ctx->w_hasumi, &ctx->w_overlap, &ctx->w_slab, &ctx->w_cob, &ctx->w_coo,
                     &ctx->w_scan, &ctx->w_err, &ctx->w_tiles, &ctx->w_vr, &ctx->w_fs, &ctx->w_gtiles, &ctx->w_gcount, &ctx->w_fstiles, &ctx->w_thdr2, &ctx->w_need, &ctx->w_rb_list, &ctx->w_rb_count, &ctx->w_stats, &ctx->w_sq_count, &ctx->w_sq_words, &ctx->w_sq_index,  &ctx->d_pair_off, &ctx->d_cref, &ctx->d_cflags, &ctx->d_umi,
                     &ctx->d_reads, &ctx->d_cigar, &ctx->d_payload, &ctx->d_pair_group, &ctx->d_ngroups, &ctx->d_groups,
                     &ctx->d_out, &ctx->d_out_bytes, &ctx->s_tid, &ctx->s_pos, &ctx->s_len, &ctx->s_off, &ctx->s_depth, &ctx->u_names, &ctx->u_off, &ctx->u_out, &ctx->u_status, &ctx->f_text, &ctx->f_anchor, &ctx->f_cnt, &ctx->f_hpos, &ctx->f_hbase,
                     &ctx->f_flag, &ctx->f_coff, &ctx->f_out};
    for (DevBuf *b : all) release(*b);
    for (int k = 0; k < GCB_MAX_CHUNKS; k++) {
        if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
        if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
        if (ctx->ev_out[k]) cudaEventDestroy(ctx->ev_out[k]);
    }
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
    if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
    if (ctx->d2h_out) cudaStreamDestroy(ctx->d2h_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *gcb_last_error(const gcb_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int gcb_set_reference_device(gcb_ctx *ctx, const uint8_t *packed4_dev, int64_t packed_bytes, const int64_t *contig_off,
                             const int64_t *contig_len, int32_t n_contigs) {
    if (!ctx || n_contigs < 0 || (n_contigs > 0 && (!packed4_dev || !contig_off || !contig_len)) || packed_bytes < 0)
        return fail(ctx, GCB_ERR_ARG, "gcb_set_reference_device: bad argument");
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->genome = {nullptr, nullptr, nullptr, 0, 0};
    if (n_contigs == 0) return GCB_OK;
    int rc;
    if ((rc = reserve(ctx, ctx->g_off, (size_t)n_contigs * 8)) != GCB_OK) return rc;
    if ((rc = reserve(ctx, ctx->g_len, (size_t)n_contigs * 8)) != GCB_OK) return rc;
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->g_off.p, contig_off, (size_t)n_contigs * 8, cudaMemcpyHostToDevice, ctx->stream));
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->g_len.p, contig_len, (size_t)n_contigs * 8, cudaMemcpyHostToDevice, ctx->stream));
    GCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->genome.packed4 = packed4_dev;
    ctx->genome.contig_off = (const int64_t *)ctx->g_off.p;
    ctx->genome.contig_len = (const int64_t *)ctx->g_len.p;
    ctx->genome.n_contigs = n_contigs;
    ctx->genome.packed_bytes = packed_bytes;
    return GCB_OK;
}

int gcb_set_reference(gcb_ctx *ctx, const uint8_t *packed4, int64_t packed_bytes, const int64_t *contig_off, const int64_t *contig_len,
                      int32_t n_contigs) {
    if (!ctx || n_contigs < 0 || (n_contigs > 0 && !packed4) || packed_bytes < 0)
        return fail(ctx, GCB_ERR_ARG, "gcb_set_reference: bad argument");
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_contigs == 0) {
        ctx->genome = {nullptr, nullptr, nullptr, 0, 0};
        return GCB_OK;
    }
    int rc;
    if ((rc = reserve(ctx, ctx->g_packed, (size_t)packed_bytes)) != GCB_OK) return rc;
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->g_packed.p, packed4, (size_t)packed_bytes, cudaMemcpyHostToDevice, ctx->stream));
    return gcb_set_reference_device(ctx, (const uint8_t *)ctx->g_packed.p, packed_bytes, contig_off, contig_len, n_contigs);
}

int gcb_consensus_batch_device(gcb_ctx *ctx, const gcb_batch *batch, gcb_result *result, uint32_t stages, void *stream_) {
    if (!ctx || !batch || !result) return fail(ctx, GCB_ERR_ARG, "gcb_consensus_batch_device: null argument");
    if (batch->n_clusters < 0 || batch->n_pairs < 0 || batch->umi_words < 1 || batch->umi_words > GCB_MAX_UMI_WORDS ||
        batch->payload_bytes < 0 || (batch->payload_bytes & 15) || ((uintptr_t)batch->payload & 15) || ((uintptr_t)result->out_payload & 3))
        return fail(ctx, GCB_ERR_ARG, "gcb_consensus_batch_device: bad sizes or alignment (payload 16 B, payload_bytes % 16, out_payload 4 B)");
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    const TilePlan plan = plan_tiles(batch->max_cluster_bytes, ctx->ring_window_shift);
    const int64_t n_tiles = (batch->payload_bytes + plan.window - 1) / plan.window;
    Workspace ws;
    int rc = reserve_workspace(ctx, batch->n_pairs, batch->n_clusters, n_tiles, batch->payload_bytes, ws);
    if (rc != GCB_OK) return rc;
    const ViewRange whole = {0, batch->n_clusters, 0, batch->n_pairs, 0, batch->payload_bytes, 0, 0, 0};
    if (stages & GCB_STAGE_UMI_GROUP) GCB_CUDA(ctx, cudaMemsetAsync(ws.error_flag, 0, 4, stream));
    return launch_stages(ctx, *batch, *result, ws, plan, whole, stages, stream, nullptr, result->out_bytes);
}

int gcb_batch_status(gcb_ctx *ctx, void *stream_) {
    if (!ctx) return GCB_ERR_ARG;
    if (!ctx->w_err.p) return GCB_OK;
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    int32_t flag = 0;
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    GCB_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->w_err.p, 4, cudaMemcpyDeviceToHost, stream));
    GCB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (flag != GCB_OK) fail(ctx, flag, flag == GCB_ERR_CAPACITY ? "out_payload too small" : "malformed batch");
    return flag;
}

static int consensus_batch_host(gcb_ctx *ctx, const gcb_batch *hb, gcb_result *hr);

int gcb_consensus_batch(gcb_ctx *ctx, const gcb_batch *hb, gcb_result *hr) {
    const int rc = consensus_batch_host(ctx, hb, hr);
    if (rc != GCB_OK && ctx) {
        // copies that read or write the caller's buffers may still be in flight on the three streams: the caller is free to
        // release them as soon as it sees the error
        cudaStreamSynchronize(ctx->h2d);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->d2h);
        cudaStreamSynchronize(ctx->d2h_out);
    }
    return rc;
}

static int consensus_batch_host(gcb_ctx *ctx, const gcb_batch *hb, gcb_result *hr) {
    if (!ctx || !hb || !hr) return fail(ctx, GCB_ERR_ARG, "gcb_consensus_batch: null argument");
    if (hb->n_clusters < 0 || hb->n_pairs < 0 || hb->payload_bytes < 0 || hb->n_cigar_ops < 0 || hr->out_capacity < 0 ||
        hb->umi_words < 1 || hb->umi_words > GCB_MAX_UMI_WORDS || (hb->payload_bytes & 15))
        return fail(ctx, GCB_ERR_ARG, "gcb_consensus_batch: bad size");
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nc = (size_t)hb->n_clusters, np = (size_t)hb->n_pairs;
    int rc;
    // the host itself walks cluster_pair_off (chunk boundaries, copy sizes): it must be monotone inside [0, n_pairs] before
    // anything is sized from it; everything else about the batch is validated on the device (GCB_ERR_MALFORMED)
    if (nc > 0) {
        if (!hb->cluster_pair_off || !hb->reads || !hb->payload) return fail(ctx, GCB_ERR_ARG, "gcb_consensus_batch: null array");
        int32_t prev = hb->cluster_pair_off[0];
        bool ok = prev >= 0;
        for (size_t c = 1; c <= nc && ok; c++) {
            const int32_t cur = hb->cluster_pair_off[c];
            ok = cur >= prev;
            prev = cur;
        }
        if (!ok || prev > hb->n_pairs) return fail(ctx, GCB_ERR_MALFORMED, "gcb_consensus_batch: cluster_pair_off is not monotone inside [0, n_pairs]");
    }
#define GCB_RESERVE(buf, bytes) if ((rc = reserve(ctx, ctx->buf, (bytes))) != GCB_OK) return rc
    GCB_RESERVE(d_pair_off, (nc + 1) * 4);
    GCB_RESERVE(d_cref, nc * 4);
    GCB_RESERVE(d_cflags, nc);
    GCB_RESERVE(d_umi, np * (size_t)hb->umi_words * 8);
    GCB_RESERVE(d_reads, 2 * np * sizeof(gcb_read_desc));
    GCB_RESERVE(d_cigar, (size_t)hb->n_cigar_ops * 4);
    GCB_RESERVE(d_payload, (size_t)hb->payload_bytes);
    GCB_RESERVE(d_pair_group, np * 4);
    GCB_RESERVE(d_ngroups, nc * 4);
    GCB_RESERVE(d_groups, np * sizeof(gcb_group_result));
    GCB_RESERVE(d_out, (size_t)hr->out_capacity);
    GCB_RESERVE(d_out_bytes, 8 * (GCB_MAX_CHUNKS + 1));
#undef GCB_RESERVE
    gcb_batch db = *hb;
    db.cluster_pair_off = (const int32_t *)ctx->d_pair_off.p;
    db.cluster_ref = (const int32_t *)ctx->d_cref.p;
    db.cluster_flags = (const uint8_t *)ctx->d_cflags.p;
    db.umi = (const uint64_t *)ctx->d_umi.p;
    db.reads = (const gcb_read_desc *)ctx->d_reads.p;
    db.cigar = (const uint32_t *)ctx->d_cigar.p;
    db.payload = (const uint8_t *)ctx->d_payload.p;
    gcb_result dr;
    dr.pair_group = (int32_t *)ctx->d_pair_group.p;
    dr.cluster_n_groups = (int32_t *)ctx->d_ngroups.p;
    dr.groups = (gcb_group_result *)ctx->d_groups.p;
    dr.out_payload = (uint8_t *)ctx->d_out.p;
    dr.out_capacity = hr->out_capacity;
    dr.out_bytes = (int64_t *)ctx->d_out_bytes.p;
    int64_t *d_totals = (int64_t *)ctx->d_out_bytes.p;  // [k]: consensus bytes emitted by chunks 0..k

    // ---- chunks of clusters of about GCB_CHUNK_BYTES of payload each: while chunk k is voted, chunk k+1 is on its
    // way in and chunk k-1 on its way out (three streams; PCIe is full duplex)
    auto slab_start = [&](int32_t c) -> int64_t {  // same rule as umi_group_kernel
        const int32_t p = hb->cluster_pair_off[c];
        return p < hb->n_pairs ? hb->reads[2 * (int64_t)p].data_off : hb->payload_bytes;
    };
    int K = (int)((hb->payload_bytes + ctx->chunk_bytes - 1) / ctx->chunk_bytes);
    if (K > GCB_MAX_CHUNKS) K = GCB_MAX_CHUNKS;
    if (K < 1) K = 1;
    if ((int64_t)K > (int64_t)nc) K = nc > 0 ? (int)nc : 1;
    const TilePlan plan = plan_tiles(hb->max_cluster_bytes, ctx->ring_window_shift);
    ViewRange view[GCB_MAX_CHUNKS];
    {
        int32_t c_prev = 0;
        int64_t tiles = 0, scans = 0;
        for (int k = 0; k < K; k++) {
            int32_t c_end = (int32_t)nc;
            if (k + 1 < K) {  // first cluster whose slab starts at or after the target byte
                const int64_t target = hb->payload_bytes / K * (k + 1);
                int32_t lo = c_prev, hi = (int32_t)nc;
                while (lo < hi) {
                    const int32_t mid = lo + (hi - lo) / 2;
                    if (slab_start(mid) < target) lo = mid + 1;
                    else hi = mid;
                }
                c_end = lo;
            }
            ViewRange &v = view[k];
            v.c0 = c_prev;
            v.c1 = c_end;
            v.p0 = nc ? hb->cluster_pair_off[v.c0] : 0;
            v.p1 = nc ? hb->cluster_pair_off[v.c1] : 0;
            v.s0 = v.c0 < (int32_t)nc ? slab_start(v.c0) : hb->payload_bytes;
            v.s1 = v.c1 < (int32_t)nc ? slab_start(v.c1) : hb->payload_bytes;
            if (k == 0) v.s0 = 0;
            v.tile_base = tiles;
            v.scan_base = scans;
            v.index = k;
            tiles += (v.s1 - v.s0 + plan.window - 1) / plan.window + 1;
            scans += (v.c1 - v.c0 + SCAN_BLOCK - 1) / SCAN_BLOCK + 1;
            c_prev = c_end;
        }
    }
    Workspace ws;
    if ((rc = reserve_workspace(ctx, hb->n_pairs, hb->n_clusters, (hb->payload_bytes + plan.window - 1) / plan.window, hb->payload_bytes, ws)) != GCB_OK) return rc;
    cudaStream_t sc = ctx->stream, sin = ctx->h2d, sout = ctx->d2h;
    struct timespec ts0, ts1, ts2, ts3;
    if (ctx->trace_e2e) {
        if (!ctx->ev_t0) { cudaEventCreate(&ctx->ev_t0); cudaEventCreate(&ctx->ev_t1); }
        clock_gettime(CLOCK_MONOTONIC, &ts0);
        cudaEventRecord(ctx->ev_t0, sin);
    }
    GCB_CUDA(ctx, cudaMemsetAsync(ws.error_flag, 0, 4, sc));
    for (int k = 0; k < K; k++) {
        const ViewRange &v = view[k];
        const size_t vc = (size_t)(v.c1 - v.c0), vp = (size_t)(v.p1 - v.p0);
#define GCB_UP(buf, src, off, bytes, elem)                                                                                   \
    if ((bytes) > 0)                                                                                                         \
    GCB_CUDA(ctx, cudaMemcpyAsync((char *)ctx->buf.p + (size_t)(off) * (elem), (const char *)(src) + (size_t)(off) * (elem), (bytes), \
                                  cudaMemcpyHostToDevice, sin))
        if (k == 0) GCB_UP(d_cigar, hb->cigar, 0, (size_t)hb->n_cigar_ops * 4, 4);
        GCB_UP(d_pair_off, hb->cluster_pair_off, v.c0, (vc + 1) * 4, 4);
        GCB_UP(d_cref, hb->cluster_ref, v.c0, vc * 4, 4);
        GCB_UP(d_cflags, hb->cluster_flags, v.c0, vc, 1);
        GCB_UP(d_umi, hb->umi, (size_t)v.p0 * hb->umi_words, vp * hb->umi_words * 8, 8);
        GCB_UP(d_reads, hb->reads, 2 * (size_t)v.p0, 2 * vp * sizeof(gcb_read_desc), sizeof(gcb_read_desc));
        GCB_UP(d_payload, hb->payload, v.s0, (size_t)(v.s1 - v.s0), 1);
#undef GCB_UP
        GCB_CUDA(ctx, cudaEventRecord(ctx->ev_in[k], sin));
        GCB_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_in[k], 0));
        if ((rc = launch_stages(ctx, db, dr, ws, plan, v, GCB_STAGE_ALL, sc, k ? d_totals + (k - 1) : nullptr, d_totals + k)) != GCB_OK) return rc;
        GCB_CUDA(ctx, cudaEventRecord(ctx->ev_done[k], sc));
        GCB_CUDA(ctx, cudaStreamWaitEvent(sout, ctx->ev_done[k], 0));
        if (vp > 0) {
            GCB_CUDA(ctx, cudaMemcpyAsync(hr->pair_group + v.p0, dr.pair_group + v.p0, vp * 4, cudaMemcpyDeviceToHost, sout));
            GCB_CUDA(ctx, cudaMemcpyAsync(hr->groups + v.p0, dr.groups + v.p0, vp * sizeof(gcb_group_result), cudaMemcpyDeviceToHost, sout));
        }
        if (vc > 0) GCB_CUDA(ctx, cudaMemcpyAsync(hr->cluster_n_groups + v.c0, dr.cluster_n_groups + v.c0, vc * 4, cudaMemcpyDeviceToHost, sout));
        GCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_totals + k, d_totals + k, 8, cudaMemcpyDeviceToHost, sout));
        GCB_CUDA(ctx, cudaEventRecord(ctx->ev_out[k], sout));
    }
    if (ctx->trace_e2e) {
        cudaEventRecord(ctx->ev_t1, sin);
        clock_gettime(CLOCK_MONOTONIC, &ts1);
    }
    // consensus records: the size of every chunk's share is known only on the device.  They leave on a stream of their own: on
    // `sout` they would queue behind the result rows of EVERY chunk (enqueued above) and all 57 MB per million pairs would
    // cross after the last chunk (measured: 1.15 ms of 11.5).  ev_out[k] implies the chunk's kernels are done.
    cudaStream_t srec = ctx->d2h_out;
    int64_t done = 0;
    for (int k = 0; k < K; k++) {
        GCB_CUDA(ctx, cudaEventSynchronize(ctx->ev_out[k]));
        int64_t upto = ctx->h_totals[k];
        if (upto > hr->out_capacity) upto = hr->out_capacity;  // (the error flag below reports it)
        if (upto > done) GCB_CUDA(ctx, cudaMemcpyAsync(hr->out_payload + done, dr.out_payload + done, (size_t)(upto - done), cudaMemcpyDeviceToHost, srec));
        if (upto > done) done = upto;
    }
    if (ctx->trace_e2e) clock_gettime(CLOCK_MONOTONIC, &ts2);
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_flag, ws.error_flag, 4, cudaMemcpyDeviceToHost, sout));
    GCB_CUDA(ctx, cudaStreamSynchronize(sout));
    GCB_CUDA(ctx, cudaStreamSynchronize(srec));
    GCB_CUDA(ctx, cudaStreamSynchronize(sc));
    if (ctx->trace_e2e) {
        clock_gettime(CLOCK_MONOTONIC, &ts3);
        float h2d_ms = 0.f;
        cudaEventElapsedTime(&h2d_ms, ctx->ev_t0, ctx->ev_t1);
        auto ms = [](const timespec &a, const timespec &b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6; };
        fprintf(stderr, "gcb_consensus_batch: %d chunks; enqueue %.3f ms, last chunk's results at %.3f ms, done at %.3f ms; the copy-in stream was busy %.3f ms\n",
                K, ms(ts0, ts1), ms(ts0, ts2), ms(ts0, ts3), (double)h2d_ms);
    }
    *hr->out_bytes = ctx->h_totals[K - 1];
    const int32_t flag = *ctx->h_flag;
    if (flag != GCB_OK) return fail(ctx, flag, flag == GCB_ERR_CAPACITY ? "out_payload too small" : "malformed batch");
    return GCB_OK;
}

int gcb_extract_umi(gcb_ctx *ctx, const char *names, const int64_t *name_off, int32_t n, const char *prefix, int32_t umi_words,
                    uint64_t *out_umi, uint8_t *status) {
    if (!ctx || n < 0 || umi_words < 1 || umi_words > GCB_MAX_UMI_WORDS || (n > 0 && (!names || !name_off || !out_umi || !status)))
        return fail(ctx, GCB_ERR_ARG, "gcb_extract_umi: bad argument");
    UmiPrefix pf;
    memset(&pf, 0, sizeof pf);
    const size_t plen = prefix ? strlen(prefix) : 0;
    if (plen >= UMI_MAX_PREFIX) return fail(ctx, GCB_ERR_ARG, "gcb_extract_umi: prefix longer than 31 characters");
    if (plen) memcpy(pf.s, prefix, plen);
    pf.len = (int32_t)plen;
    if (n == 0) return GCB_OK;
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)name_off[n];
    int rc;
    if ((rc = reserve(ctx, ctx->u_names, bytes)) != GCB_OK || (rc = reserve(ctx, ctx->u_off, ((size_t)n + 1) * 8)) != GCB_OK ||
        (rc = reserve(ctx, ctx->u_out, (size_t)n * umi_words * 8)) != GCB_OK || (rc = reserve(ctx, ctx->u_status, (size_t)n)) != GCB_OK)
        return rc;
    cudaStream_t st = ctx->stream;
    if (bytes) GCB_CUDA(ctx, cudaMemcpyAsync(ctx->u_names.p, names, bytes, cudaMemcpyHostToDevice, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->u_off.p, name_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    GCB_LAUNCH(umi_extract_kernel, dim3((unsigned)((n + UMI_EXTRACT_THREADS - 1) / UMI_EXTRACT_THREADS)), dim3(UMI_EXTRACT_THREADS), 0, st,
               (const char *)ctx->u_names.p, (const int64_t *)ctx->u_off.p, n, pf, umi_words, (uint64_t *)ctx->u_out.p, (uint8_t *)ctx->u_status.p);
    ctx->launches++;
    GCB_CUDA(ctx, cudaGetLastError());
    GCB_CUDA(ctx, cudaMemcpyAsync(out_umi, ctx->u_out.p, (size_t)n * umi_words * 8, cudaMemcpyDeviceToHost, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(status, ctx->u_status.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    GCB_CUDA(ctx, cudaStreamSynchronize(st));
    return GCB_OK;
}

int gcb_pack_fasta(gcb_ctx *ctx, const char *text, int64_t n, int32_t max_contigs, uint8_t *packed4_out, int64_t packed_cap, int64_t *contig_off,
                   int64_t *contig_len, int64_t *name_off, int32_t *name_len, int32_t *n_contigs, int64_t *packed_bytes) {
    if (!ctx || n < 0 || (n > 0 && !text) || max_contigs < 0 || !n_contigs || !packed_bytes || packed_cap < 0 ||
        (max_contigs > 0 && (!contig_off || !contig_len || !name_off || !name_len)))
        return fail(ctx, GCB_ERR_ARG, "gcb_pack_fasta: bad argument");
    *n_contigs = 0;
    *packed_bytes = 0;
    // fastareader.cpp:33-40: the reader starts behind the first '>' of the file, wherever it is
    const char *first = n > 0 ? (const char *)memchr(text, '>', (size_t)n) : nullptr;
    if (!first) return GCB_OK;
    const int64_t skip = first - text, m = n - skip;
    const int64_t n_blocks = (m + FA_BLOCK - 1) / FA_BLOCK;
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = reserve(ctx, ctx->f_text, (size_t)m)) != GCB_OK || (rc = reserve(ctx, ctx->f_anchor, (size_t)(n_blocks + 1) * 8)) != GCB_OK ||
        (rc = reserve(ctx, ctx->f_cnt, (size_t)(n_blocks + 1) * 16)) != GCB_OK || (rc = reserve(ctx, ctx->f_hpos, (size_t)max_contigs * 8)) != GCB_OK ||
        (rc = reserve(ctx, ctx->f_hbase, (size_t)max_contigs * 8)) != GCB_OK || (rc = reserve(ctx, ctx->f_flag, 4)) != GCB_OK ||
        (rc = reserve(ctx, ctx->f_coff, (size_t)max_contigs * 8)) != GCB_OK)
        return rc;
    cudaStream_t st = ctx->stream;
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->f_text.p, first, (size_t)m, cudaMemcpyHostToDevice, st));
    GCB_CUDA(ctx, cudaMemsetAsync(ctx->f_flag.p, 0, 4, st));
    const FaView v = {(const uint8_t *)ctx->f_text.p, m};
    int64_t *anchor = (int64_t *)ctx->f_anchor.p, *cnt = (int64_t *)ctx->f_cnt.p, *hpos = (int64_t *)ctx->f_hpos.p, *hbase = (int64_t *)ctx->f_hbase.p;
    GCB_LAUNCH(fa_anchor_kernel, dim3((unsigned)n_blocks), dim3(FA_THREADS), 0, st, v, anchor);
    GCB_LAUNCH(fa_scan_kernel<FaMax>, dim3(1), dim3(FA_THREADS), 0, st, anchor, n_blocks, 1);
    GCB_LAUNCH(fa_count_kernel, dim3((unsigned)n_blocks), dim3(FA_THREADS), 0, st, v, (const int64_t *)anchor, cnt);
    GCB_LAUNCH(fa_scan_kernel<FaSum>, dim3(1), dim3(FA_THREADS), 0, st, cnt, n_blocks, 2);
    GCB_LAUNCH(fa_header_kernel, dim3((unsigned)n_blocks), dim3(FA_THREADS), 0, st, v, (const int64_t *)anchor, (const int64_t *)cnt, hpos, hbase,
               max_contigs, (int32_t *)ctx->f_flag.p);
    ctx->launches += 5;
    GCB_CUDA(ctx, cudaGetLastError());
    int64_t totals[2] = {0, 0};
    int32_t flag = 0;
    GCB_CUDA(ctx, cudaMemcpyAsync(totals, cnt + 2 * n_blocks, 16, cudaMemcpyDeviceToHost, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->f_flag.p, 4, cudaMemcpyDeviceToHost, st));
    GCB_CUDA(ctx, cudaStreamSynchronize(st));
    if (flag & 2) return fail(ctx, GCB_ERR_MALFORMED, "gcb_pack_fasta: a header line the reference does not read as one ('>' before a line end or another '>')");
    if ((flag & 1) || totals[1] > max_contigs) return fail(ctx, GCB_ERR_CAPACITY, "gcb_pack_fasta: more contigs than max_contigs");
    const int32_t nc = (int32_t)totals[1];
    if (nc > 0) {
        GCB_CUDA(ctx, cudaMemcpyAsync(name_off, hpos, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
        GCB_CUDA(ctx, cudaMemcpyAsync(contig_len, hbase, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
        GCB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    int64_t bytes = 0;
    for (int32_t i = 0; i < nc; i++) {
        const int64_t end = i + 1 < nc ? contig_len[i + 1] : totals[0];
        contig_len[i] = end - contig_len[i];  // bases between this header and the next
        contig_off[i] = bytes;
        bytes += ((contig_len[i] + 1) / 2 + 15) & ~(int64_t)15;
        // the contig's id: the header up to the first space (fastareader.cpp:98-102); offsets are into the caller's text
        const int64_t h = skip + name_off[i] + 1;
        int64_t e = h;
        while (e < n && text[e] != '\n' && text[e] != ' ') e++;
        name_off[i] = h;
        name_len[i] = (int32_t)(e - h);
    }
    if (bytes == 0) bytes = 16;
    if (bytes > packed_cap || !packed4_out) return fail(ctx, GCB_ERR_CAPACITY, "gcb_pack_fasta: packed4_out too small");
    if ((rc = reserve(ctx, ctx->f_out, (size_t)bytes)) != GCB_OK) return rc;
    GCB_CUDA(ctx, cudaMemsetAsync(ctx->f_out.p, 0, (size_t)bytes, st));
    if (nc > 0) {
        GCB_CUDA(ctx, cudaMemcpyAsync(ctx->f_coff.p, contig_off, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
        GCB_LAUNCH(fa_pack_kernel, dim3((unsigned)n_blocks), dim3(FA_THREADS), 0, st, v, (const int64_t *)anchor, (const int64_t *)cnt, (const int64_t *)hbase,
                   (const int64_t *)ctx->f_coff.p, nc, (uint32_t *)ctx->f_out.p);
        ctx->launches++;
        GCB_CUDA(ctx, cudaGetLastError());
    }
    GCB_CUDA(ctx, cudaMemcpyAsync(packed4_out, ctx->f_out.p, (size_t)bytes, cudaMemcpyDeviceToHost, st));
    GCB_CUDA(ctx, cudaStreamSynchronize(st));
    *n_contigs = nc;
    *packed_bytes = bytes;
    return GCB_OK;
}

int gcb_set_chunk_bytes(gcb_ctx *ctx, int64_t bytes) {
    if (!ctx || bytes < 16) return GCB_ERR_ARG;
    ctx->chunk_bytes = bytes;
    return GCB_OK;
}

void *gcb_host_alloc(size_t bytes, int write_combined) {
    void *p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaHostAlloc(&p, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void gcb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int gcb_set_debug(gcb_ctx *ctx, int key, int value) {
    if (!ctx) return GCB_ERR_ARG;
    if (key == 2) ctx->ring_window_shift = (value == 14 || value == 15) ? value : 0;  // tuning only: same results
    else if (key == 3) ctx->group_lanes = value;                                      // tuning only: same results
    else if (key == 5) ctx->force_generic = value != 0;                               // tests: the generic kernel votes every tile
    else if (key == 6) ctx->trace_e2e = value != 0;                                   // measurement: gcb_consensus_batch's timeline on stderr
    else return GCB_ERR_ARG;
    return GCB_OK;
}

int gcb_set_slow_queue_bytes(gcb_ctx *ctx, int64_t bytes) {
    if (!ctx || bytes < 0) return GCB_ERR_ARG;
    ctx->slow_queue_bytes = bytes;
    return GCB_OK;
}

int gcb_stat_depth(gcb_ctx *ctx, const int32_t *tid, const int32_t *pos, const int32_t *l_qseq, int64_t n, int32_t coverage_step,
                   const int64_t *target_len, int32_t n_targets, int64_t *depth) {
    if (!ctx || n < 0 || coverage_step <= 0 || n_targets < 0 || (n > 0 && (!tid || !pos || !l_qseq)) || (n_targets > 0 && (!target_len || !depth)))
        return fail(ctx, GCB_ERR_ARG, "gcb_stat_depth: bad argument");
    if (n == 0 || n_targets == 0) return GCB_OK;
    std::vector<int64_t> off((size_t)n_targets + 1, 0);
    for (int32_t t = 0; t < n_targets; t++) {
        if (target_len[t] < 0) return fail(ctx, GCB_ERR_ARG, "gcb_stat_depth: negative target length");
        off[(size_t)t + 1] = off[(size_t)t] + 1 + target_len[t] / coverage_step;  // stats.cpp:43
    }
    const size_t bins = (size_t)off[(size_t)n_targets];
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = reserve(ctx, ctx->s_tid, (size_t)n * 4)) != GCB_OK || (rc = reserve(ctx, ctx->s_pos, (size_t)n * 4)) != GCB_OK ||
        (rc = reserve(ctx, ctx->s_len, (size_t)n * 4)) != GCB_OK || (rc = reserve(ctx, ctx->s_off, off.size() * 8)) != GCB_OK ||
        (rc = reserve(ctx, ctx->s_depth, bins * 8)) != GCB_OK)
        return rc;
    cudaStream_t st = ctx->stream;
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->s_tid.p, tid, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->s_pos.p, pos, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->s_len.p, l_qseq, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->s_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
    GCB_CUDA(ctx, cudaMemcpyAsync(ctx->s_depth.p, depth, bins * 8, cudaMemcpyHostToDevice, st));
    GCB_LAUNCH(stat_depth_kernel, dim3((unsigned)((n + DEPTH_THREADS - 1) / DEPTH_THREADS)), dim3(DEPTH_THREADS), 0, st, (const int32_t *)ctx->s_tid.p,
               (const int32_t *)ctx->s_pos.p, (const int32_t *)ctx->s_len.p, n, coverage_step, (const int64_t *)ctx->s_off.p, n_targets,
               (unsigned long long *)ctx->s_depth.p);
    ctx->launches++;
    GCB_CUDA(ctx, cudaGetLastError());
    GCB_CUDA(ctx, cudaMemcpyAsync(depth, ctx->s_depth.p, bins * 8, cudaMemcpyDeviceToHost, st));
    GCB_CUDA(ctx, cudaStreamSynchronize(st));
    return GCB_OK;
}

int gcb_get_cluster_stats(gcb_ctx *ctx, gcb_cluster_stats *out, int reset) {
    if (!ctx || !out) return GCB_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->w_stats.p) return GCB_OK;  // nothing processed yet
    GCB_CUDA(ctx, cudaSetDevice(ctx->device));
    GCB_CUDA(ctx, cudaMemcpyAsync(out, ctx->w_stats.p, sizeof *out, cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) GCB_CUDA(ctx, cudaMemsetAsync(ctx->w_stats.p, 0, sizeof *out, ctx->stream));
    GCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GCB_OK;
}

int64_t gcb_launch_count(const gcb_ctx *ctx) { return ctx ? ctx->launches : 0; }

#ifdef GCB_SIMT_CHECK
// tests only (CPU SIMT-check build): which vote path the tiles and columns took
void gcb_simt_counters(int64_t *out, int reset) {
    for (int k = 0; k < 8; k++) {
        out[k] = g_simt_counters[k];
        if (reset) g_simt_counters[k] = 0;
    }
}
// tests only: Cluster::isDuplex on two encoded UMIs, by the literal split (bit 0) and by the strand forms duplex_kernel compares (bit 1)
int gcb_simt_is_duplex(const uint64_t *a, const uint64_t *b, int nw) {
    const Umi ua = umi_load(a, nw), ub = umi_load(b, nw);
    Umi ca, sa, cb, sb;
    const bool ta = umi_strand_forms(ua, ca, sa), tb = umi_strand_forms(ub, cb, sb);
    return (umi_is_duplex(ua, ub) ? 1 : 0) | ((ta && tb && umi_equal(ca, sb)) ? 2 : 0);
}
#endif

}  // extern "C"
