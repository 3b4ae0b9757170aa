// k_score_vote.cuh — the GENERIC vote kernel: Pair::computeScore (pair.cpp:88-172) fused with
// Group::makeConsensus (group.cpp:320-579) for the tiles the tiled kernel (k_vote_tiled.cuh) cannot
// take: clusters larger than its staging buffer, tiles with more pairs than its tables hold, fields
// that overflow its 16-bit table slots.  It has no size limits at all.
//
// Work decomposition
//   CTA   = loops over the tiles listed in ws.generic_tiles.  A tile's clusters are one contiguous
//           byte range of the payload; runs of clusters that fit the staging buffer are staged
//           into shared memory with a single bulk asynchronous copy (cp.async.bulk -> UBLKCP)
//           that completes on an mbarrier.  A cluster larger than the staging buffer is voted
//           straight from global memory by the same code (generic pointers).
//   warp  = one cluster at a time (dynamic counter); inside it one (family, side) after another.
//   lane  = one column of the template per pass of 32 columns.
// Scores are never materialised: each (column, read) recomputes the overlap score and the rewritten
// quality (pair.cpp:158-159) of its base from the two mates' bytes in shared memory.
#pragma once

#include "k_group_select.cuh"
#include "vote_column.cuh"

namespace gcb {

constexpr int VOTE_THREADS = 128;
constexpr int SLAB_CAP = 40 * 1024;     // staging buffer; tile overflow is handled by sub-tiling
constexpr int VOTE_SMEM = SLAB_CAP + 64;

// ---- bulk asynchronous copy global -> shared, completion on an mbarrier ------------------------------
#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_barrier_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: arm the barrier with the byte count, then start the copy
__device__ __forceinline__ void tile_load_async(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tile_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GCB_TILE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GCB_TILE_DONE;\n"
        "bra GCB_TILE_WAIT;\n"
        "GCB_TILE_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#else
inline void tile_barrier_init(uint64_t *) {}
inline void tile_load_async(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
inline void tile_wait(uint64_t *, uint32_t) {}
#endif

// ---- what a lane needs to know about one read of the family ----------------------------------------
struct ReadInfo {
    int own_off;   // record offset relative to the cluster's slab
    int own_l;
    int shift;     // readpos = column + shift (group.cpp:377-379)
    int mate_off;
    int mate_l;
    int ov_own;    // overlap window start in this read / in its mate / length (pair.cpp:108-119)
    int ov_mate;
    int ov_len;
    int flags;     // RI_*
};
constexpr int RI_VOTES = 1, RI_OVERLAP = 2, RI_TEMPLATE = 4;

GCB_DEV ReadInfo read_info_shfl(const ReadInfo &v, int src) {
    ReadInfo r;
    r.own_off = __shfl_sync(FULL, v.own_off, src);
    r.own_l = __shfl_sync(FULL, v.own_l, src);
    r.shift = __shfl_sync(FULL, v.shift, src);
    r.mate_off = __shfl_sync(FULL, v.mate_off, src);
    r.mate_l = __shfl_sync(FULL, v.mate_l, src);
    r.ov_own = __shfl_sync(FULL, v.ov_own, src);
    r.ov_mate = __shfl_sync(FULL, v.ov_mate, src);
    r.ov_len = __shfl_sync(FULL, v.ov_len, src);
    r.flags = __shfl_sync(FULL, v.flags, src);
    return r;
}

struct SideCtx {
    const uint8_t *cbase;  // the cluster's slab (shared or global memory)
    int64_t slab0;         // payload offset of the slab
    int side;
    int mb, m;             // the family's pairs: ws.members[mb .. mb+m)
    int tmpl;              // template read slot
    int l_out;
    bool left_mode;
};

GCB_DEV ReadInfo make_read_info(const BatchView &b, const Workspace &ws, const SideCtx &s, int pair) {
    ReadInfo ri;
    const int own = 2 * pair + s.side, mate = 2 * pair + (1 - s.side);
    const gcb_read_desc od = b.reads[own], md = b.reads[mate];
    ri.flags = 0;
    ri.own_off = 0; ri.own_l = 0; ri.shift = 0; ri.mate_off = 0; ri.mate_l = 0; ri.ov_own = 0; ri.ov_mate = 0; ri.ov_len = 0;
    if (od.l_qseq < 0) return ri;
    const uint8_t f = ws.vote_flags[own];
    if (!(f & VOTE_PARTICIPATES)) return ri;
    ri.flags = RI_VOTES | (own == s.tmpl ? RI_TEMPLATE : 0);
    ri.own_off = (int)(od.data_off - s.slab0);
    ri.own_l = od.l_qseq;
    const int d = (f & VOTE_LENDIFF0) ? 0 : od.l_qseq - s.l_out;  // group.cpp:339-349
    ri.shift = s.left_mode ? 0 : d;
    const PairOverlap ov = ws.overlap[pair];
    if (ov.valid) {
        ri.flags |= RI_OVERLAP;
        ri.mate_off = (int)(md.data_off - s.slab0);
        ri.mate_l = md.l_qseq;
        ri.ov_own = s.side == 0 ? ov.left_start : ov.right_start;
        ri.ov_mate = s.side == 0 ? ov.right_start : ov.left_start;
        ri.ov_len = ov.cmp_len;
    }
    return ri;
}

// base, rewritten quality and score of read `ri` at template column i (pair.cpp:88-172 for one base)
GCB_DEV bool fetch_base(const uint8_t *cbase, const ReadInfo &ri, int i, int side, const gcb_options &o, int &base, int &qual, int &score) {
    const int rp = i + ri.shift;
    if (rp < 0 || rp >= ri.own_l) return false;
    const uint8_t *q = cbase + ri.own_off;
    const uint8_t *sq = q + GCB_ALIGN4(ri.own_l);
    qual = q[rp];
    base = base_at(sq, rp);
    if (!(ri.flags & RI_OVERLAP)) {  // pair.cpp:92,99: no mate or no M block: the moderate score everywhere
        score = sc8(o.score_moderate);
        return true;
    }
    const int k = rp - ri.ov_own;
    if (k < 0 || k >= ri.ov_len) {  // pair.cpp:121-131
        score = qual2score(o, qual);
        return true;
    }
    const int mp = ri.ov_mate + k;
    if (mp < 0 || mp >= ri.mate_l) {  // cannot happen for a CIGAR consistent with l_qseq
        score = sc8(o.score_moderate);
        return true;
    }
    const uint8_t *mq = cbase + ri.mate_off;
    const int mqual = mq[mp];
    const int mbase = base_at(mq + GCB_ALIGN4(ri.mate_l), mp);
    if (base == mbase) {  // pair.cpp:147-152
        score = sc8(qual2score(o, (qual + mqual) / 2) + 4);
    } else {  // pair.cpp:153-169
        const int lq = side == 0 ? qual : mqual, rq = side == 0 ? mqual : qual;
        const bool left_wins = lq >= rq;
        const bool mine = side == 0 ? left_wins : !left_wins;
        score = mine ? sc8(qual2score(o, lq >= rq ? lq - rq : rq - lq) - 3) : 0;
        qual = max(0, qual - mqual);
    }
    return true;
}

// Calls f(base, qual, score, is_template) for every voting read of the family at this lane's column.
// All 32 lanes must call it (it shuffles); `active` lanes are the ones whose column exists.
template <typename F>
GCB_DEV void for_each_vote(const BatchView &b, const Workspace &ws, const gcb_options &o, const SideCtx &s, ReadInfo &cached,
                           bool &cached_valid, int i, bool active, F &&f) {
    const int lane = lane_id();
    for (int chunk = 0; chunk < s.m; chunk += WARP) {
        if (!(cached_valid && s.m <= WARP)) {
            const int k = chunk + lane;
            if (k < s.m) cached = make_read_info(b, ws, s, ws.members[s.mb + k]);
            else cached.flags = 0;
            cached_valid = true;
        }
        const int cnt = min(WARP, s.m - chunk);
        for (int k = 0; k < cnt; k++) {
            const int fl = __shfl_sync(FULL, cached.flags, k);
            if (!(fl & RI_VOTES)) continue;
            const ReadInfo ri = read_info_shfl(cached, k);
            int base, qual, score;
            if (active && fetch_base(s.cbase, ri, i, s.side, o, base, qual, score)) f(base, qual, score, (ri.flags & RI_TEMPLATE) != 0);
        }
    }
}

// up to three distinct codes of a column in registers; a fourth raises `overflow`
struct SparseBins {
    VoteBin s[3];
    int total;
    bool overflow;
    GCB_DEV void init() {
        s[0].base = s[1].base = s[2].base = -1;
        total = 0;
        overflow = false;
    }
    GCB_DEV void add(int base, int qual, int score) {
        total += score;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (s[k].base == base) {
                s[k].cnt++;
                s[k].score += score;
                s[k].qual += qual;
                s[k].maxq = max(s[k].maxq, qual);
                return;
            }
            if (s[k].base < 0) {
                s[k].base = base;
                s[k].cnt = 1;
                s[k].score = score;
                s[k].qual = qual;
                s[k].maxq = qual;
                return;
            }
        }
        overflow = true;
    }
};

// pack this pass's 32 qualities / 32 nibbles and store them as aligned words
GCB_DEV void store_pass(uint8_t *out, int l_out, int col0, int qual, int nib) {
    const int lane = lane_id();
    const int i = col0 + lane;
    unsigned w = (unsigned)qual & 0xFFu;
    w |= __shfl_down_sync(FULL, w, 1) << 8;
    w |= __shfl_down_sync(FULL, w, 2) << 16;
    if ((lane & 3) == 0 && i < GCB_ALIGN4(l_out)) *(uint32_t *)(out + i) = w;
    unsigned byte = ((unsigned)nib & 0xFu) << 4;
    byte |= __shfl_down_sync(FULL, (unsigned)nib & 0xFu, 1);
    unsigned v = byte & 0xFFu;
    v |= (__shfl_down_sync(FULL, v, 2) & 0xFFu) << 8;
    v |= (__shfl_down_sync(FULL, v, 4) & 0xFFFFu) << 16;
    if ((lane & 7) == 0 && (i >> 1) < GCB_ALIGN4((l_out + 1) >> 1)) *(uint32_t *)(out + GCB_ALIGN4(l_out) + (i >> 1)) = v;
}

// One (family, side): group.cpp:320-579.  Whole warp.
// `abs_off`: out_off already holds the absolute offset (a tile that a tile-preparing vote kernel accepted and handed over later)
GCB_DEV void vote_family_side(const BatchView &b, const ResultView &r, const Workspace &ws, const GenomeView &gv, const gcb_options &o,
                              const uint8_t *cbase, int64_t slab0, int c, int slot, int side, int64_t out_base, bool abs_off) {
    const int lane = lane_id();
    const uint8_t mode = ws.side_mode[2 * (int64_t)slot + side];
    if (mode == SIDE_NONE) return;
    gcb_group_result *gr = r.groups + slot;
    const int tmpl = gr->tmpl_read[side];
    const int64_t out_off = abs_off ? gr->out_off[side] : out_base + gr->out_off[side];
    __syncwarp();
    if (lane == 0) gr->out_off[side] = out_off;
    const gcb_read_desc od = b.reads[tmpl];
    const int l_out = od.l_qseq;
    const int64_t rec = record_bytes(l_out);
    if (out_off + rec > r.out_capacity) {
        if (lane == 0) raise_error(ws.error_flag, GCB_ERR_CAPACITY);
        return;
    }
    uint8_t *out = r.out_payload + out_off;
    const uint8_t *trec = cbase + (od.data_off - slab0);

    if (mode == SIDE_COPY) {  // group.cpp:73-77: the record itself, padding bytes zeroed
        const int nc = max(GCB_ALIGN4(l_out), 2 * GCB_ALIGN4((l_out + 1) >> 1));
        for (int col0 = 0; col0 < nc; col0 += WARP) {
            const int i = col0 + lane;
            const int q = i < l_out ? trec[i] : 0;
            const int nb = i < 2 * ((l_out + 1) >> 1) ? base_at(trec + GCB_ALIGN4(l_out), i) : 0;
            store_pass(out, l_out, col0, q, nb);
        }
        return;
    }

    SideCtx s;
    s.cbase = cbase;
    s.slab0 = slab0;
    s.side = side;
    s.mb = ws.group_off[slot];
    {
        const int G = r.cluster_n_groups[c], p0 = b.cluster_pair_off[c];
        const int me = (slot - p0) + 1 < G ? ws.group_off[slot + 1] : b.cluster_pair_off[c + 1];
        s.m = me - s.mb;
    }
    s.tmpl = tmpl;
    s.l_out = l_out;
    s.left_mode = mode == SIDE_LEFT;
    const uint32_t *ocig = b.cigar + od.cigar_off;
    const int ncig = od.n_cigar;

    ReadInfo cached = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool cached_valid = false;

    // group.cpp:354-360: a template without CIGAR votes only over the shortest read's length
    int len = l_out;
    if (ncig == 0) {
        int mn = l_out;
        for (int k = lane; k < s.m; k += WARP) {
            const int sl = 2 * ws.members[s.mb + k] + side;
            if (b.reads[sl].l_qseq >= 0 && (ws.vote_flags[sl] & VOTE_PARTICIPATES)) mn = min(mn, b.reads[sl].l_qseq);
        }
        len = warp_min(mn);
    }
    // group.cpp:362-367 + reference.cpp:33-71
    const uint8_t *refdata = nullptr;
    if (od.isize != 0) {
        const int contig = b.cluster_ref[c];
        if (gv.packed4 && contig >= 0 && contig < gv.n_contigs) {
            const int64_t span = (int64_t)get_ref_offset(ocig, ncig, len - 1) + 1;
            if ((int64_t)od.pos + span < gv.contig_len[contig]) refdata = gv.packed4 + gv.contig_off[contig];
        }
    }
    const int64_t contig_len = refdata ? gv.contig_len[b.cluster_ref[c]] : 0;

    // the template as this side's vote sees it (its own rewritten qualities)
    ReadInfo tinfo = make_read_info(b, ws, s, tmpl >> 1);

    int diff = 0, mm_inc = 0;
    const int ncols = max(GCB_ALIGN4(l_out), 2 * GCB_ALIGN4((l_out + 1) >> 1));
    for (int col0 = 0; col0 < ncols; col0 += WARP) {
        const int i = col0 + lane;
        const bool active = i < len;
        // what the record holds at this column before the vote (also the result beyond `len`)
        int obase = 0, oqual = 0;
        if (i < l_out) {
            int sc;
            fetch_base(cbase, tinfo, i, side, o, obase, oqual, sc);
        } else if (i < 2 * ((l_out + 1) >> 1)) {
            obase = base_at(trec + GCB_ALIGN4(l_out), i);  // the unused low nibble of an odd-length record
        }
        SparseBins bins;
        bins.init();
        for_each_vote(b, ws, o, s, cached, cached_valid, i, active,
                      [&](int base, int qual, int score, bool) { bins.add(base, qual, score); });

        VoteBin obs[16];
        int nobs = 0;
        int total = bins.total;
        if (__any_sync(FULL, active && bins.overflow)) {  // four or more distinct codes in some column: full histogram
            int cnt[16], scs[16], qls[16], mxq[16];
            for (int k = 0; k < 16; k++) cnt[k] = scs[k] = qls[k] = mxq[k] = 0;
            int tot = 0;
            for_each_vote(b, ws, o, s, cached, cached_valid, i, active, [&](int base, int qual, int score, bool) {
                cnt[base]++;
                scs[base] += score;
                qls[base] += qual;
                mxq[base] = max(mxq[base], qual);
                tot += score;
            });
            if (bins.overflow) {
                total = tot;
                for (int k = 0; k < 16; k++)
                    if (cnt[k] > 0) {
                        obs[nobs].base = k; obs[nobs].cnt = cnt[k]; obs[nobs].score = scs[k]; obs[nobs].qual = qls[k]; obs[nobs].maxq = mxq[k];
                        nobs++;
                    }
            }
        }
        if (!bins.overflow) {
            for (int k = 0; k < 3; k++)
                if (bins.s[k].base >= 0) obs[nobs++] = bins.s[k];
        }

        int new_base = obase, new_qual = oqual;
        bool want_rescan = false;
        ColumnTop top;
        int ref4 = 0;
        bool decided = false;
        if (active) {
            top = column_top(o, obs, nobs, total);
            if (top.fast) {
                new_qual = top.top.maxq;  // group.cpp:422-426: the base is NOT written
                decided = true;
            } else {
                if (refdata) {  // group.cpp:430-439
                    const int refpos = get_ref_offset(ocig, ncig, i);
                    if (refpos >= 0) {
                        const int64_t gp = (int64_t)od.pos + refpos;
                        if (gp < contig_len) {
                            const uint8_t two = refdata[gp >> 1];
                            ref4 = genome_nibble_to_bam((gp & 1) ? (two >> 4) : (two & 0xF));
                        }
                    }
                }
                if (top.need_ref && ref4 != 0) {
                    int rmax = 0;
                    for (int k = 0; k < nobs; k++)
                        if (obs[k].base == ref4) rmax = obs[k].maxq;
                    if (rmax >= 128) want_rescan = true;  // `char refBaseQual` wraps: order matters (group.cpp:474-490)
                }
            }
        }
        int rbq = 0;
        bool any_high = false;
        if (__any_sync(FULL, want_rescan)) {
            // sequential restatement of group.cpp:474-490, template first, then the family in map order
            int tb, tq, ts;
            if (active && fetch_base(cbase, tinfo, i, side, o, tb, tq, ts) && tb == ref4) {
                if (tq > rbq) rbq = sc8(tq);
                if (tq >= o.high_quality) any_high = true;
            }
            for_each_vote(b, ws, o, s, cached, cached_valid, i, active, [&](int base, int qual, int, bool is_tmpl) {
                if (is_tmpl || base != ref4) return;
                if (qual > rbq) rbq = sc8(qual);
                if (qual >= o.high_quality) any_high = true;
            });
        }
        if (active && !decided) {
            if (!want_rescan) {
                rbq = 0;
                any_high = false;
                for (int k = 0; k < nobs; k++)
                    if (obs[k].base == ref4) { rbq = obs[k].maxq; any_high = obs[k].maxq >= o.high_quality; }
            }
            const ColumnOut co = column_arbitrate(o, top, ref4, rbq, any_high);
            if (obase != co.base) {  // group.cpp:509-524
                new_base = co.base;
                diff++;
                if (ref4 != 0) {
                    if (obase == ref4) mm_inc++;
                    else if (co.base == ref4) mm_inc--;
                }
            }
            new_qual = co.qual;
        }
        if (i >= l_out) new_qual = 0;
        store_pass(out, l_out, col0, new_qual, new_base);
    }
    diff = warp_sum(diff);
    mm_inc = warp_sum(mm_inc);
    if (mm_inc > 5) {  // group.cpp:538-566: put the (score-rewritten) template back
        for (int col0 = 0; col0 < ncols; col0 += WARP) {
            const int i = col0 + lane;
            int obase = 0, oqual = 0, sc;
            if (i < l_out) fetch_base(cbase, tinfo, i, side, o, obase, oqual, sc);
            else if (i < 2 * ((l_out + 1) >> 1)) obase = base_at(trec + GCB_ALIGN4(l_out), i);
            store_pass(out, l_out, col0, oqual, obase);
        }
    }
    if (lane == 0) {
        gr->diff[side] = diff;
        gr->mismatch_inc[side] = mm_inc;
    }
}

__global__ void __launch_bounds__(VOTE_THREADS) score_vote_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o) {
    GCB_GRID_DEP();
    GCB_DYN_SMEM(smem);
    if (batch_is_malformed(ws.error_flag)) return;
    uint64_t *bar = (uint64_t *)smem;
    int *counter = (int *)(smem + 8);
    int *s_ce = (int *)(smem + 12);
    uint8_t *slab = smem + 64;
    const int lane = lane_id();
    const int tid = (int)threadIdx.x;
    if (tid == 0) tile_barrier_init(bar);
    __syncthreads();
    uint32_t parity = 0;
    const int n_generic = *ws.generic_count;
    for (int gi = (int)blockIdx.x; gi < n_generic; gi += (int)gridDim.x) {
    const int tile_code = ws.generic_tiles[gi];  // ~tile: the consensus offsets of the tile's families are already absolute
    const bool abs_off = tile_code < 0;
    const int tile = abs_off ? ~tile_code : tile_code;
    const int c0 = ws.tile_dir[tile].c0, c1 = ws.tile_dir[tile + 1].c0;
    int cs = c0;
    while (cs < c1) {
        // the longest run of clusters starting at cs whose slabs fit the staging buffer
        const int64_t start = ws.slab_off[cs];
        int ce;
        if (ws.slab_off[c1] - start <= SLAB_CAP) {
            ce = c1;
        } else {
            if (tid == 0) *s_ce = cs;
            __syncthreads();
            for (int cc = cs + 1 + tid; cc <= c1; cc += VOTE_THREADS)
                if (ws.slab_off[cc] - start <= SLAB_CAP) atomicMax(s_ce, cc);
            __syncthreads();
            ce = *s_ce;
        }
        const bool staged = ce > cs;
        if (!staged) ce = cs + 1;  // one cluster bigger than the buffer: voted from global memory
        if (tid == 0) {
            *counter = cs;
            if (staged) tile_load_async(slab, b.payload + start, (uint32_t)(ws.slab_off[ce] - start), bar);
        }
        __syncthreads();
        if (staged) {
            tile_wait(bar, parity);
            parity ^= 1;
        }
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(counter, 1);
            c = __shfl_sync(FULL, c, 0);
            if (c >= ce) break;
            const int64_t slab0 = ws.slab_off[c];
            const uint8_t *cbase = staged ? slab + (slab0 - start) : b.payload + slab0;
            const int64_t out_base = ws.scan_block[c / SCAN_BLOCK] + ws.cluster_out_off[c];
            const int G = r.cluster_n_groups[c], p0 = b.cluster_pair_off[c];
            for (int g = 0; g < G; g++)
                for (int side = 0; side < 2; side++) vote_family_side(b, r, ws, gv, o, cbase, slab0, c, p0 + g, side, out_base, abs_off);
        }
        __syncthreads();
        cs = ce;
    }
    }
}

}  // namespace gcb
