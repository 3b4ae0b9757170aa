// k_group_select.cuh — the two descriptor-only kernels that run before the vote:
//   umi_group_kernel        Cluster::clusterByUMI's grouping loop        cluster.cpp:55-100
//   select_template_kernel  Group::consensusMerge / consensusMergeBam    group.cpp:68-318 (everything but makeConsensus)
// Both give one warp to one cluster; they read UMIs, read descriptors and CIGARs only (the payload
// is touched only by the >1000-pair low-complexity probe).
#pragma once

#include "device_common.cuh"

namespace gcb {

// coverage counters of the CPU SIMT-check build (tests only): ring tiles, generic tiles, voted columns, slow columns,
// uniform family sides, non-uniform family sides, clusters selected in registers, arena wrap-arounds
#ifdef GCB_SIMT_CHECK
inline int64_t g_simt_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define GCB_COUNT(k, n) (g_simt_counters[k] += (n))
#define GCB_TRACE(tag) ::simt::trace(tag)
#else
#define GCB_COUNT(k, n) ((void)0)
#define GCB_TRACE(tag) ((void)0)
#endif

struct BatchView {  // gcb_batch with device pointers
    int32_t n_clusters, n_pairs, umi_words;
    const int32_t *cluster_pair_off;
    const int32_t *cluster_ref;
    const uint8_t *cluster_flags;
    const uint64_t *umi;
    const gcb_read_desc *reads;
    const uint32_t *cigar;
    const uint8_t *payload;
    int64_t payload_bytes;   // end of this view's payload (absolute offset)
    int64_t payload_origin;  // start of this view's payload: vote tiles are windows counted from here
    int64_t n_cigar_ops;     // entries of `cigar`
};

// The documented preconditions of a gcb_batch are checked on the device, where the descriptors are read anyway: a batch that
// breaks one raises GCB_ERR_MALFORMED, the offending cluster emits nothing, and every later kernel of the batch returns at once
// (a misplaced slab would otherwise reach the bulk copies of the vote as an unaligned address: a sticky fault, not an error code).
GCB_DEV bool batch_is_malformed(const int32_t *error_flag) { return *(volatile const int32_t *)error_flag == GCB_ERR_MALFORMED; }
// a read's record must lie inside its cluster's slab [lo, hi) at a 4-byte boundary, its CIGAR inside the pool
GCB_DEV bool read_desc_ok(const BatchView &b, const gcb_read_desc &d, int64_t lo, int64_t hi) {
    if (d.l_qseq < 0) return true;  // an empty slot
    return (d.data_off & 3) == 0 && d.data_off >= lo && d.data_off + record_bytes(d.l_qseq) <= hi && d.cigar_off >= 0 &&
           (int64_t)d.cigar_off + d.n_cigar <= b.n_cigar_ops;
}

struct ResultView {  // gcb_result with device pointers
    int32_t *pair_group;
    int32_t *cluster_n_groups;
    gcb_group_result *groups;
    uint8_t *out_payload;
    int64_t out_capacity;
    int64_t *out_bytes;
};

constexpr int GROUP_THREADS = 128;  // 4 clusters per CTA
#ifndef GCB_SELECT_MINB
#define GCB_SELECT_MINB 12  // CTAs per SM select_template_kernel is compiled for (42 registers a thread)
#endif

// ------------------------------------------------------------------------------------------------
// UMIs of NW 64-bit words held in registers (the kernel is instantiated for 1, 2 and GCB_MAX_UMI_WORDS words:
// sixteen characters fit one word, a duplex UMI of 8+1+8 needs two)
template <int NW>
struct UmiT {
    uint64_t w[NW];
};
template <int NW>
GCB_DEV UmiT<NW> umit_load(const uint64_t *p) {
    UmiT<NW> u;
#pragma unroll
    for (int k = 0; k < NW; k++) u.w[k] = p[k];
    return u;
}
template <int NW>
GCB_DEV bool umit_equal(const UmiT<NW> &a, const UmiT<NW> &b) {
    bool e = true;
#pragma unroll
    for (int k = 0; k < NW; k++) e = e && a.w[k] == b.w[k];
    return e;
}
template <int NW>
GCB_DEV bool umit_less(const UmiT<NW> &a, const UmiT<NW> &b) {  // std::string order of the decoded UMIs
#pragma unroll
    for (int k = 0; k < NW; k++)
        if (a.w[k] != b.w[k]) return a.w[k] < b.w[k];
    return false;
}
template <int NW>
GCB_DEV int umit_diff(const UmiT<NW> &a, const UmiT<NW> &b) {  // Cluster::umiDiff, cluster.cpp:41-53
    int d = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) d += nibble_diff64(a.w[k], b.w[k]);
    return d;
}

// cluster.cpp:55-100.  umiCount (a map<string,int>) becomes a per-pair multiplicity; each round takes
// the lexicographically first UMI of maximal count among the pairs still unassigned and absorbs every
// unassigned pair within `thr` of it, in map (= pair) order.  Counts never need decrementing: pairs
// carrying the same UMI are always absorbed together.  `window_shift`: the vote tile window is 1 << window_shift.
template <int NW, int GS>
__global__ void __launch_bounds__(GROUP_THREADS) umi_group_kernel(BatchView b, ResultView r, Workspace ws, int32_t window_shift,
                                                                   int32_t n_tiles) {
    GCB_GRID_DEP();
    const Grp<GS> g;
    const int lane = g.gl;
    const int c = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (WARP / GS) + (lane_id() / GS);
    if (c >= b.n_clusters) return;
    const int p0 = b.cluster_pair_off[c], p1 = b.cluster_pair_off[c + 1], n = p1 - p0;
    const int thr = b.cluster_flags[c] >> GCB_CLUSTER_UMI_THR_SHIFT;
    const uint64_t *umi = b.umi + (int64_t)p0 * NW;
    // a cluster of at most GS pairs is grouped in registers, a pair per lane: its UMIs are requested here, together with the
    // descriptors the checks below read (the kernel waits for loads, not for instructions)
    UmiT<NW> u;
    {
        const bool pre = lane < n && n <= GS && p0 >= 0 && p1 <= b.n_pairs;
#pragma unroll
        for (int k = 0; k < NW; k++) u.w[k] = pre ? umi[(int64_t)lane * NW + k] : 0ull;
    }
    {   // preconditions: monotone pair offsets inside the batch; the cluster's slab 16-byte aligned, after the previous cluster's,
        // inside this view's payload
        bool bad = p0 < 0 || p1 < p0 || p1 > b.n_pairs;
        if (!bad) {
            const int64_t s = p0 < b.n_pairs ? b.reads[2 * (int64_t)p0].data_off : b.payload_bytes;
            bad = (s & 15) != 0 || s < b.payload_origin || s > b.payload_bytes;
            if (!bad && c > 0) {
                const int pp = b.cluster_pair_off[c - 1];
                bad = pp < 0 || pp > p0 || (pp < b.n_pairs && b.reads[2 * (int64_t)pp].data_off > s);
            }
        }
        if (bad) {  // (uniform over the cluster's lanes)
            if (lane == 0) {
                raise_error(ws.error_flag, GCB_ERR_MALFORMED);
                r.cluster_n_groups[c] = 0;
            }
            return;
        }
    }

    if (lane == 0) {
        // slab bounds of this cluster and the vote kernel's tile directory: tile t owns the clusters
        // whose slab starts inside [t*window, (t+1)*window)
        const int64_t s = p0 < b.n_pairs ? b.reads[2 * (int64_t)p0].data_off : b.payload_bytes;
        ws.slab_off[c] = s;
        int64_t t_lo = 0;
        if (c > 0) {
            const int pp = b.cluster_pair_off[c - 1];
            const int64_t prev = pp < b.n_pairs ? b.reads[2 * (int64_t)pp].data_off : b.payload_bytes;
            t_lo = ((prev - b.payload_origin) >> window_shift) + 1;
        }
        const TileDir here = {c, p0, s};
        for (int64_t t = t_lo; t <= ((s - b.payload_origin) >> window_shift) && t <= n_tiles; t++) ws.tile_dir[t] = here;
        if (c == b.n_clusters - 1) {
            ws.slab_off[c + 1] = b.payload_bytes;
            const TileDir end = {b.n_clusters, p1, b.payload_bytes};
            for (int64_t t = ((s - b.payload_origin) >> window_shift) + 1; t <= n_tiles; t++) ws.tile_dir[t] = end;
        }
    }

    if (n <= GS) {
        // The usual cluster in registers, a pair per lane: the multiplicity of a pair's UMI is the population of a match mask,
        // the round's top UMI comes from word-wise minimum reductions, the unassigned pairs are a bit mask.  (The loops below
        // re-read every UMI of the cluster from memory for every pair and every round.)
        const bool act = lane < n;
        const unsigned actmask = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
        unsigned eq = actmask;
#pragma unroll
        for (int k = 0; k < NW; k++) eq &= __match_any_sync(g.mask, u.w[k]) >> g.base;
        const int cnt = __popc(eq);  // cluster.cpp:57-65
        const bool has_umi = g.any(act && (u.w[0] >> 60) != 0);
        if (lane == 0) ws.cluster_has_umi[c] = has_umi ? 1 : 0;
        unsigned un = actmask;  // the pairs no family has absorbed yet
        int gi = 0, filled = 0, mine = -1;
        while (un != 0u) {  // cluster.cpp:66-100
            const bool cand = ((un >> lane) & 1u) != 0u;
            const int top = g.max_of(cand ? cnt : -1);
            bool in = cand && cnt == top;  // of the most frequent UMIs, the first in string order
            UmiT<NW> best;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                const unsigned hi = (unsigned)(u.w[k] >> 32), lo = (unsigned)u.w[k];
                const unsigned mh = __reduce_min_sync(g.mask, in ? hi : 0xFFFFFFFFu);
                in = in && hi == mh;
                const unsigned ml = __reduce_min_sync(g.mask, in ? lo : 0xFFFFFFFFu);
                in = in && lo == ml;
                best.w[k] = ((uint64_t)mh << 32) | ml;
            }
            const bool absorb = cand && umit_diff<NW>(u, best) <= thr;
            const unsigned m = g.ballot(absorb);
            if (absorb) {
                ws.members[p0 + filled + __popc(m & ((1u << lane) - 1u))] = p0 + lane;
                mine = gi;
            }
            if (lane == 0) ws.group_off[p0 + gi] = p0 + filled;
            if (m == 0u) {  // cannot happen (the top UMI is within 0 of itself); never spin on bad input
                if (lane == 0) raise_error(ws.error_flag, GCB_ERR_MALFORMED);
                break;
            }
            filled += __popc(m);
            un &= ~m;
            gi++;
        }
        if (act) r.pair_group[p0 + lane] = mine;
        if (lane == 0) r.cluster_n_groups[c] = gi;
        return;
    }

    // multiplicity of every pair's UMI inside the cluster (cluster.cpp:57-65)
    bool has = false;
    for (int i = lane; i < n; i += GS) {
        const UmiT<NW> u = umit_load<NW>(umi + (int64_t)i * NW);
        has |= (u.w[0] >> 60) != 0;
        int cnt = 0;
        for (int j = 0; j < n; j++) cnt += umit_equal<NW>(u, umit_load<NW>(umi + (int64_t)j * NW)) ? 1 : 0;
        ws.scratch[2 * (int64_t)(p0 + i)] = cnt;
        r.pair_group[p0 + i] = -1;
    }
    has = g.any(has);
    if (lane == 0) ws.cluster_has_umi[c] = has ? 1 : 0;
    g.sync();  // (the rounds below read pair_group / scratch entries written by other lanes of the group)

    int filled = 0, gi = 0;
    while (filled < n) {  // cluster.cpp:66-100
        int best_cnt = -1;
        UmiT<NW> best;
#pragma unroll
        for (int k = 0; k < NW; k++) best.w[k] = 0ull;
        for (int i = lane; i < n; i += GS) {
            if (r.pair_group[p0 + i] >= 0) continue;
            const int cnt = ws.scratch[2 * (int64_t)(p0 + i)];
            const UmiT<NW> u = umit_load<NW>(umi + (int64_t)i * NW);
            if (cnt > best_cnt || (cnt == best_cnt && umit_less<NW>(u, best))) { best_cnt = cnt; best = u; }
        }
        for (int off = GS / 2; off > 0; off >>= 1) {
            const int oc = g.shfl_xor(best_cnt, off);
            UmiT<NW> ou;
#pragma unroll
            for (int k = 0; k < NW; k++) ou.w[k] = g.shfl_xor(best.w[k], off);
            if (oc > best_cnt || (oc == best_cnt && oc >= 0 && umit_less<NW>(ou, best))) { best_cnt = oc; best = ou; }
        }
        const int start = filled;
        for (int base = 0; base < n; base += GS) {
            const int i = base + lane;
            bool absorb = false;
            if (i < n && r.pair_group[p0 + i] < 0) absorb = umit_diff<NW>(umit_load<NW>(umi + (int64_t)i * NW), best) <= thr;
            const unsigned m = g.ballot(absorb);
            if (absorb) {
                ws.members[p0 + filled + __popc(m & ((1u << lane) - 1u))] = p0 + i;
                r.pair_group[p0 + i] = gi;
            }
            filled += __popc(m);
        }
        if (lane == 0) ws.group_off[p0 + gi] = p0 + start;
        if (filled == start) {  // cannot happen (the top UMI is within 0 of itself); never spin on bad input
            if (lane == 0) raise_error(ws.error_flag, GCB_ERR_MALFORMED);
            break;
        }
        gi++;
        g.sync();
    }
    if (lane == 0) r.cluster_n_groups[c] = gi;
}

// ------------------------------------------------------------------------------------------------
// group.cpp:136-313 for one side of one family: returns the template's read slot (or -1) and fills
// vote_flags / side_mode for the vote kernel.  Called by a whole warp; the result is warp-uniform.
struct SideChoice {
    int out;     // template read slot, -1 = NULL
    int k;       // its index in the family
    int len;     // voted columns (group.cpp:354-360)
    bool fits;   // every VoteRead field fits its 16 bits
    bool uniform;  // every voter shares the template's geometry (FS_UNIFORM)
};

// The VoteRead of read slot sk (see device_common.cuh); `fits` is cleared when a field overflows.
GCB_DEV VoteRead make_vote_read(const BatchView &b, const Workspace &ws, int64_t slab0, int sk, int side, uint8_t f, int l_out, bool left_mode,
                                bool &fits) {
    VoteRead v = {VR_NO_VOTE, 0, 0, 0, 0, 0, 0, 0};
    if (!(f & VOTE_PARTICIPATES)) return v;
    const gcb_read_desc rd = b.reads[sk];
    const int64_t off = rd.data_off - slab0;
    const int d = (f & VOTE_LENDIFF0) ? 0 : rd.l_qseq - l_out;  // group.cpp:339-349
    const int shift = left_mode ? 0 : d;
    if (off < 0 || (off >> 2) >= VR_NO_VOTE || rd.l_qseq > 0x7FFF || shift < -0x8000 || shift > 0x7FFF) {
        fits = false;
        return v;
    }
    v.own_off4 = (uint16_t)(off >> 2);
    v.own_l = (int16_t)rd.l_qseq;
    v.shift = (int16_t)shift;
    v.ov_len = VR_NO_OVERLAP_INFO;
    const PairOverlap ov = ws.overlap[sk >> 1];
    if (ov.valid) {
        const gcb_read_desc md = b.reads[sk ^ 1];
        const int64_t moff = md.data_off - slab0;
        // (32-bit sums of 32-bit CIGAR offsets and lengths: saturate instead of wrapping on absurd input)
        const int64_t own64 = side == 0 ? ov.left_start : ov.right_start, mate64 = side == 0 ? ov.right_start : ov.left_start;
        const int lim = 0x3FFFFFFF;
        int own = (int)(own64 < -lim ? -lim : own64 > lim ? lim : own64);
        int mate = (int)(mate64 < -lim ? -lim : mate64 > lim ? lim : mate64);
        int len = ov.cmp_len < -lim ? -lim : ov.cmp_len > lim ? lim : ov.cmp_len;
        const int a = own > 0 ? own : 0;
        const int e = own + len < rd.l_qseq ? own + len : rd.l_qseq;
        if (len > 0 && e > a) {  // the window clipped to this read's own indices
            mate += a - own;
            own = a;
            len = e - a;
        } else {
            own = mate = len = 0;
        }
        if (moff < 0 || (moff >> 2) > 0xFFFF || md.l_qseq > 0x7FFF || mate < -0x8000 || mate > 0x7FFF) {
            fits = false;
            return v;
        }
        v.mate_off4 = (uint16_t)(moff >> 2);
        v.mate_l = (int16_t)md.l_qseq;
        v.ov_own = (int16_t)own;
        v.ov_mate = (int16_t)mate;
        v.ov_len = (int16_t)len;
    }
    return v;
}

template <int GS>
GCB_DEV SideChoice side_select(const Grp<GS> &g, const BatchView &b, const Workspace &ws, const gcb_options &o, int mb, int m, int side, int slot,
                               int64_t slab0) {
    const int lane = g.gl;
    const SideChoice none = {-1, 0, 0, true, false};
    const bool isLeft = side == 0;
    const int thr = o.skip_low_complexity_cluster_threshold;
#define GCB_SLOT(k) (2 * ws.members[mb + (k)] + side)
#define GCB_HAVE(k) (b.reads[GCB_SLOT(k)].l_qseq >= 0)
#define GCB_CIG(s) (b.cigar + b.reads[s].cigar_off)

    if (m > thr) {  // group.cpp:142-175: many distinct CIGARs + a low-complexity first read => skip the side
        int distinct = 0, first_k = 0x7FFFFFFF;
        for (int k = lane; k < m; k += GS) {
            if (!GCB_HAVE(k)) continue;
            first_k = min(first_k, k);
            const int sk = GCB_SLOT(k);
            bool seen = false;
            for (int j = 0; j < k && !seen; j++) {
                if (!GCB_HAVE(j)) continue;
                const int sj = GCB_SLOT(j);
                seen = same_cigar_string(GCB_CIG(sj), b.reads[sj].n_cigar, GCB_CIG(sk), b.reads[sk].n_cigar);
            }
            if (!seen) distinct++;
        }
        distinct = g.sum(distinct);
        first_k = g.min_of(first_k);
        if ((double)distinct > m * 0.1 && first_k != 0x7FFFFFFF) {
            const gcb_read_desc rd = b.reads[GCB_SLOT(first_k)];
            const uint8_t *seq = b.payload + rd.data_off + GCB_ALIGN4(rd.l_qseq);
            int dn = 0;
            for (int i = lane; i < rd.l_qseq - 1; i += GS)
                if (base_letter(base_at(seq, i)) != base_letter(base_at(seq, i + 1))) dn++;
            dn = g.sum(dn);
            if ((double)dn < rd.l_qseq * 0.5) return none;
        }
    }

    // Shortcut for the usual family: every read present, one identical CIGAR op, same length, same position.
    // Then every read is part of every other (bamutil.cpp:204-255 on equal CIGARs), the counts tie at m, the lengths
    // tie, so the template is the first read in map order, every read votes, and right reads share their position
    // (left-aligned columns).  Exactly what the general code below computes, without its O(m^2) CIGAR walks.
    bool same = m <= thr;
    {
        const gcb_read_desc r0 = b.reads[GCB_SLOT(0)];
        same = same && r0.l_qseq >= 0 && r0.n_cigar == 1;
        const uint32_t c0w = same ? b.cigar[r0.cigar_off] : 0u;
        for (int k = lane; k < m && same; k += GS) {
            const gcb_read_desc rk = b.reads[GCB_SLOT(k)];
            same = rk.l_qseq == r0.l_qseq && rk.n_cigar == 1 && rk.pos == r0.pos && b.cigar[rk.cigar_off] == c0w;
        }
        same = g.all(same);
    }
    bool leftReadMode = true;
    int best_cnt = m, best_k = 0;
    if (!same) {
        leftReadMode = isLeft;
        if (!isLeft) {  // group.cpp:177-194: every right read starts at the same position => left-aligned columns
            int lo = 0x7FFFFFFF, hi = -0x7FFFFFFF;
            for (int k = lane; k < m; k += GS) {
                if (!GCB_HAVE(k)) continue;
                const int pos = b.reads[GCB_SLOT(k)].pos;
                lo = min(lo, pos);
                hi = max(hi, pos);
            }
            lo = g.min_of(lo);
            hi = g.max_of(hi);
            if (lo >= hi) leftReadMode = true;  // all equal, or no read at all
        }

        // BamUtil::getRightRefPos (bamutil.cpp:379-383) of the right reads: only this general path compares them
        if (!isLeft) {
            for (int k = lane; k < m; k += GS) {
                if (!GCB_HAVE(k)) continue;
                const gcb_read_desc rk = b.reads[GCB_SLOT(k)];
                ws.right_ref_pos[GCB_SLOT(k)] = rk.pos < 0 ? -1 : rk.pos + cigar_ref_len(b.cigar + rk.cigar_off, rk.n_cigar);
            }
            g.sync();
        }
        // group.cpp:196-233: containedBy[i] = 1 + #{j : read i is part of read j}
        int first_big = 0x7FFFFFFF;
        for (int k = lane; k < m; k += GS) {
            int cnt = 0;
            if (GCB_HAVE(k)) {
                cnt = 1;
                const int si = GCB_SLOT(k);
                const uint32_t *ci = GCB_CIG(si);
                const int ni = b.reads[si].n_cigar;
                const int rrp = ws.right_ref_pos[si];
                for (int j = 0; j < m; j++) {
                    if (j == k || !GCB_HAVE(j)) continue;
                    const int sj = GCB_SLOT(j);
                    if (!isLeft && rrp != ws.right_ref_pos[sj]) continue;
                    if (is_part_of(ci, ni, GCB_CIG(sj), b.reads[sj].n_cigar, leftReadMode)) cnt++;
                }
                if (m > thr && cnt >= m / 2) first_big = min(first_big, k);
            }
            ws.scratch[2 * (int64_t)(mb + k) + side] = cnt;
        }
        first_big = g.min_of(first_big);  // group.cpp:231-232: the scan stops there, later entries stay 0

        // group.cpp:235-261: most contained, ties -> strictly shorter read, else first in map order
        int best_len = 0;
        best_cnt = -1;
        best_k = 0x7FFFFFFF;
        for (int k = lane; k < m; k += GS) {
            const int cnt = k > first_big ? 0 : ws.scratch[2 * (int64_t)(mb + k) + side];
            const int len = GCB_HAVE(k) ? b.reads[GCB_SLOT(k)].l_qseq : 0;
            if (cnt > best_cnt || (cnt == best_cnt && len < best_len)) { best_cnt = cnt; best_len = len; best_k = k; }
        }
        for (int off = GS / 2; off > 0; off >>= 1) {
            const int oc = g.shfl_xor(best_cnt, off), ol = g.shfl_xor(best_len, off),
                      ok = g.shfl_xor(best_k, off);
            if (oc > best_cnt || (oc == best_cnt && (ol < best_len || (ol == best_len && ok < best_k)))) {
                best_cnt = oc; best_len = ol; best_k = ok;
            }
        }
    }
    if ((double)best_cnt < m * 0.4 && m != 1) return none;  // group.cpp:264
    if (!GCB_HAVE(best_k)) return none;                     // group.cpp:270-285
    const int out = GCB_SLOT(best_k);
    const gcb_read_desc od = b.reads[out];

    // group.cpp:287-313 (who votes) and 339-349 (whose length difference is ignored)
    bool fits = true;
    int mn = od.l_qseq;
    VoteRead *vr = ws.vote_reads + 2 * (int64_t)mb + (int64_t)side * m;
    VoteRead mine_v = {VR_NO_VOTE, 0, 0, 0, 0, 0, 0, 0};  // the entry this lane wrote last (its only one when m <= GS)
    for (int k = lane; k < m; k += GS) {
        const VoteRead zero = {VR_NO_VOTE, 0, 0, 0, 0, 0, 0, 0};
        if (!GCB_HAVE(k)) {
            vr[k] = zero;
            mine_v = zero;
            continue;
        }
        const int sk = GCB_SLOT(k);
        uint8_t f = 0;
        if (k == best_k || same) f = VOTE_PARTICIPATES;
        else {
            const gcb_read_desc rd = b.reads[sk];
            if (is_part_of(GCB_CIG(out), od.n_cigar, GCB_CIG(sk), rd.n_cigar, leftReadMode)) {
                f = VOTE_PARTICIPATES;
                if (rd.l_qseq != od.l_qseq && rd.pos == od.pos && is_part_of(GCB_CIG(out), od.n_cigar, GCB_CIG(sk), rd.n_cigar, true))
                    f |= VOTE_LENDIFF0;
            }
        }
        ws.vote_flags[sk] = f;
        if (f & VOTE_PARTICIPATES) mn = min(mn, b.reads[sk].l_qseq);
        mine_v = make_vote_read(b, ws, slab0, sk, side, f, od.l_qseq, leftReadMode, fits);
        vr[k] = mine_v;
    }
    if (lane == 0) ws.side_mode[2 * (int64_t)slot + side] = leftReadMode ? SIDE_LEFT : SIDE_RIGHT;
    SideChoice ch;
    ch.out = out;
    ch.k = best_k;
    ch.len = od.n_cigar == 0 ? g.min_of(mn) : od.l_qseq;  // group.cpp:354-360: no CIGAR => only the shortest read's columns
    ch.fits = g.all(fits);
    // FS_UNIFORM: every voter is as long as the template, is read at the template's columns, meets its mate through
    // the same overlap window and finds its mate's record at the same distance from its own (true for every family
    // of a fixed-length library packed pair by pair)
    bool uni;
    if (m <= GS) {  // every entry is still in the register of the lane that made it: the template's comes by shuffle
        VoteRead tv;
        {
            uint32_t w[4];
            memcpy(w, &mine_v, 16);
#pragma unroll
            for (int q = 0; q < 4; q++) w[q] = __shfl_sync(g.mask, w[q], g.base + best_k);
            memcpy(&tv, w, 16);
        }
        const VoteRead v = mine_v;
        uni = ch.len == od.l_qseq && tv.shift == 0 && tv.own_l == od.l_qseq;
        if (lane < m && v.own_off4 != VR_NO_VOTE)
            uni = uni && v.own_l == tv.own_l && v.shift == 0 && v.ov_len == tv.ov_len &&
                  (v.ov_len <= 0 || (v.ov_own == tv.ov_own && v.ov_mate == tv.ov_mate && v.mate_l == tv.mate_l &&
                                     (uint16_t)(v.mate_off4 - v.own_off4) == (uint16_t)(tv.mate_off4 - tv.own_off4)));
    } else {
        g.sync();
        const VoteRead tv = vr[best_k];
        uni = ch.len == od.l_qseq && tv.shift == 0 && tv.own_l == od.l_qseq;
        for (int k = lane; k < m; k += GS) {
            const VoteRead v = vr[k];
            if (v.own_off4 == VR_NO_VOTE) continue;
            uni = uni && v.own_l == tv.own_l && v.shift == 0 && v.ov_len == tv.ov_len &&
                  (v.ov_len <= 0 || (v.ov_own == tv.ov_own && v.ov_mate == tv.ov_mate && v.mate_l == tv.mate_l &&
                                     (uint16_t)(v.mate_off4 - v.own_off4) == (uint16_t)(tv.mate_off4 - tv.own_off4)));
        }
    }
    ch.uniform = g.all(uni);
    return ch;
#undef GCB_SLOT
#undef GCB_HAVE
#undef GCB_CIG
}

// ------------------------------------------------------------------------------------------------
// The usual cluster in registers: at most GS pairs, every pair with both reads, every read one CIGAR op, and inside every
// family the reads of a side identical in length, position and CIGAR.  Then (group.cpp:136-313) every read is part of every
// other, the containment counts tie, the lengths tie, so a family's template is its first read in map order, every read
// votes, and the columns are left-aligned on both sides — what side_select computes for such a family through its `same`
// shortcut.  Here lane i of the group holds pair i's two descriptors for the whole cluster: the family of a pair comes
// from pair_group (members are in pair order), the template's geometry by shuffle, and nothing is read twice — the general
// path below walks a chain of dependent loads (members -> descriptor -> CIGAR) per family and side.
// Returns false (nothing written) when the cluster is not of that kind.
template <int GS>
GCB_DEV bool select_cluster_fast(const Grp<GS> &g, const BatchView &b, const ResultView &r, const Workspace &ws, const GenomeView &gv, int c, int p0,
                                 int n, int G) {
    const int lane = g.gl;
    const bool act = lane < n;
    const int64_t pair = p0 + (act ? lane : 0);
    const uint4 *dp = (const uint4 *)(b.reads + 2 * pair);
    const uint4 l0 = dp[0], l1 = dp[1], r0 = dp[2], r1 = dp[3];  // gcb_read_desc: {data_off lo, hi, l_qseq, pos} {isize, cigar_off, n_cigar | l_qname << 16, -}
    const int gid = act ? r.pair_group[pair] : -1;
    const int64_t slab0 = ws.slab_off[c];
    const int L_l = (int)l0.z, L_pos = (int)l0.w, L_isize = (int)l1.x, L_ncig = (int)(l1.z & 0xFFFFu), L_lqn = (int)(l1.z >> 16);
    const int R_l = (int)r0.z, R_pos = (int)r0.w, R_isize = (int)r1.x, R_ncig = (int)(r1.z & 0xFFFFu), R_lqn = (int)(r1.z >> 16);
    const int64_t L_off = ((int64_t)l0.y << 32 | l0.x) - slab0, R_off = ((int64_t)r0.y << 32 | r0.x) - slab0;
    bool ok = !act || (L_l >= 0 && R_l >= 0 && L_ncig == 1 && R_ncig == 1);
    // every field of a VoteRead must fit its 16 bits (make_vote_read)
    ok = ok && (!act || (L_off >= 0 && R_off >= 0 && (L_off >> 2) < VR_NO_VOTE && (R_off >> 2) < VR_NO_VOTE && L_l <= 0x7FFF && R_l <= 0x7FFF));
    if (!g.all(ok)) return false;
    const uint32_t L_cig = act ? b.cigar[(int)l1.y] : 0u, R_cig = act ? b.cigar[(int)r1.y] : 0u;
    // pair.cpp:103-119: the overlap window of the pair (both reads have one op: an M block or nothing)
    PairOverlap ov = {0, 0, 0, 0};
    {
        const int ll = cig_op(L_cig) == OP_MATCH ? cig_len(L_cig) : 0, rl = cig_op(R_cig) == OP_MATCH ? cig_len(R_cig) : 0;
        if (ll > 0 && rl > 0) {
            const int posDis = R_pos - L_pos;
            ov.valid = 1;
            if (posDis >= 0) {
                ov.left_start = posDis;
                ov.right_start = 0;
                ov.cmp_len = min(ll - posDis, rl);
            } else {
                ov.left_start = 0;
                ov.right_start = 0 - posDis;
                ov.cmp_len = min(ll, rl + posDis);
            }
        }
    }
    // the two VoteReads of the pair (make_vote_read with every read voting, no column shift)
    VoteRead v[2];
    bool fits = true;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        VoteRead w = {VR_NO_VOTE, 0, 0, 0, 0, 0, 0, 0};
        const int own_l = side == 0 ? L_l : R_l, mate_l = side == 0 ? R_l : L_l;
        w.own_off4 = (uint16_t)((side == 0 ? L_off : R_off) >> 2);
        w.own_l = (int16_t)own_l;
        w.shift = 0;
        w.ov_len = VR_NO_OVERLAP_INFO;
        if (ov.valid) {
            int own = side == 0 ? ov.left_start : ov.right_start, mate = side == 0 ? ov.right_start : ov.left_start, len = ov.cmp_len;
            const int lim = 0x3FFFFFFF;
            own = own < -lim ? -lim : own > lim ? lim : own;
            mate = mate < -lim ? -lim : mate > lim ? lim : mate;
            len = len < -lim ? -lim : len > lim ? lim : len;
            const int a = own > 0 ? own : 0;
            const int e = own + len < own_l ? own + len : own_l;
            if (len > 0 && e > a) {  // the window clipped to this read's own indices
                mate += a - own;
                own = a;
                len = e - a;
            } else {
                own = mate = len = 0;
            }
            if (mate < -0x8000 || mate > 0x7FFF) fits = false;
            w.mate_off4 = (uint16_t)((side == 0 ? R_off : L_off) >> 2);
            w.mate_l = (int16_t)mate_l;
            w.ov_own = (int16_t)own;
            w.ov_mate = (int16_t)mate;
            w.ov_len = (int16_t)len;
        }
        v[side] = w;
    }
    if (!g.all(!act || fits)) return false;
    // every family: its members are the lanes whose pair_group names it, in lane (= map) order.  Everything below is
    // lane-parallel but for one short loop over the families (a loop in which one lane writes a family's rows while the others
    // wait was half of this kernel's instructions).
    const unsigned fm = __match_any_sync(g.mask, gid) >> g.base;  // this lane's family (the idle lanes form one of their own)
    const int first = __ffs((int)fm) - 1, src = first + g.base;
    const int m = __popc(fm), rank = __popc(fm & ((1u << lane) - 1u));
    const bool is_first = act && lane == first;
    const unsigned firsts = g.ballot(is_first);
    // (cannot fail: umi_group_kernel numbers the families 0 .. G-1 and every family has a pair)
    if (!g.all(!act || (gid >= 0 && gid < G)) || __popc(firsts) != G) return false;
    {
        const int fl = __shfl_sync(g.mask, L_l, src), fp = __shfl_sync(g.mask, L_pos, src), gl_ = __shfl_sync(g.mask, R_l, src),
                  gp = __shfl_sync(g.mask, R_pos, src);
        const uint32_t fc = __shfl_sync(g.mask, L_cig, src), gc = __shfl_sync(g.mask, R_cig, src);
        const bool same = !act || (L_l == fl && L_pos == fp && L_cig == fc && R_l == gl_ && R_pos == gp && R_cig == gc);
        if (!g.all(same)) return false;
    }
    // FS_UNIFORM per family side ((same length, no shift: given) every member has the template's overlap window and the same
    // distance from its record to its mate's): three words per side compared with the first member's
    bool uni[2];
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const VoteRead &x = v[side];
        const bool ovl = x.ov_len > 0;
        const uint32_t k1 = (uint32_t)(uint16_t)x.ov_len | (ovl ? (uint32_t)(uint16_t)x.ov_own << 16 : 0u);
        const uint32_t k2 = ovl ? ((uint32_t)(uint16_t)x.ov_mate | (uint32_t)(uint16_t)x.mate_l << 16) : 0u;
        const uint32_t k3 = ovl ? (uint32_t)(uint16_t)(x.mate_off4 - x.own_off4) : 0u;
        const bool u = k1 == __shfl_sync(g.mask, k1, src) && k2 == __shfl_sync(g.mask, k2, src) && k3 == __shfl_sync(g.mask, k3, src);
        uni[side] = (g.ballot(act && !u) & fm) == 0u;
    }
    // where the family's members and its consensus records begin: sums over the families created before it (umi_group_kernel
    // numbers them in creation order; pairs in the high byte, record bytes — at most 32 * 2 * 48 KB — below)
    const uint32_t own_sum = is_first ? ((uint32_t)m << 24) + (uint32_t)(record_bytes(L_l) + record_bytes(R_l)) : 0u;
    uint32_t excl = 0u, total = 0u;
    for (int gi = 0; gi < G; gi++) {
        const unsigned f = g.ballot(is_first && gid == gi);  // exactly one lane
        const uint32_t sgi = __shfl_sync(g.mask, own_sum, __ffs((int)f) - 1 + g.base);
        if (gid > gi) excl += sgi;
        total += sgi;
    }
    const int mb = p0 + (int)(excl >> 24);
    const int64_t orel0 = (int64_t)(excl & 0xFFFFFFu);

    // ---- the cluster is of the usual kind: write what select_template_kernel's general path would
    if (act) {
        ws.overlap[pair] = ov;
        *(uint16_t *)(ws.vote_flags + 2 * pair) = (uint16_t)(VOTE_PARTICIPATES | (VOTE_PARTICIPATES << 8));
        ws.vote_reads[2 * (int64_t)mb + rank] = v[0];
        ws.vote_reads[2 * (int64_t)mb + m + rank] = v[1];
    }
    if (lane >= G && act) {  // slots that hold no family
        *(uint16_t *)(ws.side_mode + 2 * pair) = (uint16_t)(SIDE_NONE | (SIDE_NONE << 8));
        int2 *row = (int2 *)(r.groups + pair);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(gcb_group_result) / 8); q++) row[q] = make_int2(0, 0);
    }
    if (is_first) {  // the family's result row and its two family-side descriptors
        const int contig = b.cluster_ref[c];
        const int slot = p0 + gid;
        const int left = 2 * (int)pair, right = left + 1;
        gcb_group_result gr;
        gr.tmpl_read[0] = left;
        gr.tmpl_read[1] = right;
        gr.qname_donor[0] = gr.qname_donor[1] = -1;
        int name_slot;
        if (L_lqn <= R_lqn) { gr.qname_donor[1] = left; name_slot = left; }  // group.cpp:114-123
        else { gr.qname_donor[0] = right; name_slot = right; }
        gr.diff[0] = gr.diff[1] = 0;
        gr.mismatch_inc[0] = gr.mismatch_inc[1] = 0;
        gr.merge_reads = m;
        gr.reverse_merge_reads = 0;
        gr.status = 0;
        gr.duplex_partner = -1;
        gr.duplex_diff = 0;
        gr.umi_pair = name_slot / 2;
        int64_t orel = orel0;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const int l_out = side == 0 ? L_l : R_l, pos = side == 0 ? L_pos : R_pos, isize = side == 0 ? L_isize : R_isize;
            const uint32_t cg = side == 0 ? L_cig : R_cig;
            FsDesc fd = {0, 0, 0, 0, 0, SIDE_NONE, 0, 0, 0, 0, 0};
            fd.c = c;
            if (isize != 0 && gv.packed4 && contig >= 0 && contig < gv.n_contigs) {  // group.cpp:362-367 + reference.cpp:33-71
                const int64_t span = (int64_t)get_ref_offset(&cg, 1, l_out - 1) + 1;
                if ((int64_t)pos + span < gv.contig_len[contig]) {
                    fd.flags |= FS_REF_OK;
                    fd.ref_nib0 = 2 * gv.contig_off[contig] + pos;
                }
            }
            const int op = cig_op(cg);
            if (query_consum(op) && ref_consum(op) && cig_len(cg) >= l_out) fd.flags |= FS_SIMPLE_CIGAR;
            if (uni[side]) fd.flags |= FS_UNIFORM;
            gr.out_off[side] = orel;  // cluster-relative; the vote rebases it after the scan
            fd.mb = mb;
            fd.m = (uint16_t)m;
            fd.l_out = (uint16_t)l_out;
            fd.len = (uint16_t)l_out;
            fd.tmpl_k = 0;
            fd.mode = SIDE_LEFT;
            fd.out_rel = (uint32_t)orel;
            orel += record_bytes(l_out);
            ws.fs_desc[2 * (int64_t)slot + side] = fd;
        }
        *(uint16_t *)(ws.side_mode + 2 * (int64_t)slot) = (uint16_t)(SIDE_LEFT | (SIDE_LEFT << 8));
        r.groups[slot] = gr;
    }
    if (lane == 0) {
        ws.cluster_out_bytes[c] = (int64_t)(total & 0xFFFFFFu);
        GCB_COUNT(6, 1);
    }
    return true;
}


// group.cpp:68-134 per family of the cluster + the per-pair overlap windows of pair.cpp:103-119
template <int GS>
__global__ void __launch_bounds__(GROUP_THREADS, GCB_SELECT_MINB) select_template_kernel(BatchView b, ResultView r, Workspace ws, GenomeView gv, gcb_options o) {
    GCB_GRID_DEP();
    const Grp<GS> g;
    const int lane = g.gl;
    const int c = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (WARP / GS) + (lane_id() / GS);
    if (c >= b.n_clusters || batch_is_malformed(ws.error_flag)) return;
    const int p0 = b.cluster_pair_off[c], p1 = b.cluster_pair_off[c + 1], n = p1 - p0;
    const int G = r.cluster_n_groups[c];
    const bool crossContig = (b.cluster_flags[c] & GCB_CLUSTER_CROSS_CONTIG) != 0;
    {   // preconditions of the cluster's reads (see read_desc_ok)
        const int64_t lo = ws.slab_off[c], hi = ws.slab_off[c + 1];
        bool bad = false;
        for (int i = lane; i < n; i += GS) {
            const int64_t pair = p0 + i;
            bad = bad || !read_desc_ok(b, b.reads[2 * pair], lo, hi) || !read_desc_ok(b, b.reads[2 * pair + 1], lo, hi) || b.reads[2 * pair].l_qseq < 0;
        }
        if (g.any(bad)) {
            if (lane == 0) {
                raise_error(ws.error_flag, GCB_ERR_MALFORMED);
                ws.cluster_out_bytes[c] = 0;
            }
            return;
        }
    }
    if (n > 0 && n <= GS && !crossContig && select_cluster_fast<GS>(g, b, r, ws, gv, c, p0, n, G)) return;

    for (int i = lane; i < n; i += GS) {
        const int64_t pair = p0 + i;
        const gcb_read_desc L = b.reads[2 * pair], R = b.reads[2 * pair + 1];
        ws.vote_flags[2 * pair] = 0;
        ws.vote_flags[2 * pair + 1] = 0;
        PairOverlap ov = {0, 0, 0, 0};
        if (L.l_qseq >= 0 && R.l_qseq >= 0) {  // pair.cpp:103-119
            int lo, ll, ro, rl;
            get_m_offset_and_len(b.cigar + L.cigar_off, L.n_cigar, lo, ll);
            get_m_offset_and_len(b.cigar + R.cigar_off, R.n_cigar, ro, rl);
            if (ll > 0 && rl > 0) {
                const int posDis = R.pos - L.pos;
                ov.valid = 1;
                if (posDis >= 0) {
                    ov.left_start = lo + posDis;
                    ov.right_start = ro;
                    ov.cmp_len = min(ll - posDis, rl);
                } else {
                    ov.left_start = lo;
                    ov.right_start = ro - posDis;
                    ov.cmp_len = min(ll, rl + posDis);
                }
            }
        }
        ws.overlap[pair] = ov;
    }
    g.sync();

    const int64_t slab0 = ws.slab_off[c];
    for (int i = G + lane; i < n; i += GS) {  // slots that hold no family: the vote kernel reads side_mode to know
        ws.side_mode[2 * (int64_t)(p0 + i)] = SIDE_NONE;
        ws.side_mode[2 * (int64_t)(p0 + i) + 1] = SIDE_NONE;
        int2 *row = (int2 *)(r.groups + (p0 + i));  // and their result rows read as zeros (every other row is written below)
        static_assert(sizeof(gcb_group_result) % 8 == 0, "result rows are written as 8-byte words");
#pragma unroll
        for (int q = 0; q < (int)(sizeof(gcb_group_result) / 8); q++) row[q] = make_int2(0, 0);
    }
    int64_t out_rel = 0;
    for (int gi = 0; gi < G; gi++) {
        const int slot = p0 + gi;
        const int mb = ws.group_off[slot];
        const int me = gi + 1 < G ? ws.group_off[slot + 1] : p1;
        const int m = me - mb;
        gcb_group_result gr;
        gr.out_off[0] = gr.out_off[1] = -1;
        gr.tmpl_read[0] = gr.tmpl_read[1] = -1;
        gr.qname_donor[0] = gr.qname_donor[1] = -1;
        gr.diff[0] = gr.diff[1] = 0;
        gr.mismatch_inc[0] = gr.mismatch_inc[1] = 0;
        gr.merge_reads = m;
        gr.reverse_merge_reads = 0;
        gr.status = 0;
        gr.duplex_partner = -1;
        gr.duplex_diff = 0;
        gr.umi_pair = -1;
        const int first = ws.members[mb];
        uint8_t mode0 = SIDE_NONE;
        SideChoice ch[2] = {{-1, 0, 0, true, false}, {-1, 0, 0, true, false}};
        if (m == 1 && b.reads[2 * (int64_t)first + 1].l_qseq < 0) {  // group.cpp:73-77: passes through untouched
            gr.merge_reads = 1;
            if (b.reads[2 * (int64_t)first].l_qseq >= 0) {
                gr.tmpl_read[0] = 2 * first;
                mode0 = SIDE_COPY;
                ch[0].out = 2 * first;
                ch[0].len = b.reads[2 * (int64_t)first].l_qseq;
                const VoteRead v = make_vote_read(b, ws, slab0, 2 * first, 0, VOTE_PARTICIPATES, ch[0].len, true, ch[0].fits);
                if (lane == 0) {
                    const VoteRead zero = {VR_NO_VOTE, 0, 0, 0, 0, 0, 0, 0};
                    ws.vote_reads[2 * (int64_t)mb] = v;
                    ws.vote_reads[2 * (int64_t)mb + 1] = zero;
                }
            }
            gr.umi_pair = first;
            if (lane == 0) {
                ws.side_mode[2 * (int64_t)slot] = mode0;
                ws.side_mode[2 * (int64_t)slot + 1] = SIDE_NONE;
            }
        } else {
            int nameToCopy = -1;  // group.cpp:79-99: shortest padded qname among the left reads, first in map order
            if (crossContig) {
                long long key = 0x7FFFFFFFFFFFFFFFll;
                for (int k = lane; k < m; k += GS) {
                    const int sl = 2 * ws.members[mb + k];
                    if (b.reads[sl].l_qseq < 0) continue;
                    const long long kk = ((long long)b.reads[sl].l_qname << 32) | (unsigned)k;
                    key = kk < key ? kk : key;
                }
                for (int off = GS / 2; off > 0; off >>= 1) {
                    const long long ok = g.shfl_xor(key, off);
                    key = ok < key ? ok : key;
                }
                if (key != 0x7FFFFFFFFFFFFFFFll) nameToCopy = 2 * ws.members[mb + (int)(key & 0xFFFFFFFFll)];
            }
            if (lane == 0) {
                ws.side_mode[2 * (int64_t)slot] = SIDE_NONE;
                ws.side_mode[2 * (int64_t)slot + 1] = SIDE_NONE;
            }
            g.sync();
#pragma unroll 1  // (one copy of side_select: the kernel is 88 KB of code, and on ragged libraries it waits for instructions)
            for (int s = 0; s < 2; s++) ch[s] = side_select<GS>(g, b, ws, o, mb, m, s, slot, slab0);
            const int left = ch[0].out, right = ch[1].out;
            gr.tmpl_read[0] = left;
            gr.tmpl_read[1] = right;
            int name_slot;
            if (crossContig) {  // group.cpp:109-113
                if (left >= 0 && nameToCopy >= 0 && nameToCopy != left) gr.qname_donor[0] = nameToCopy;
                name_slot = left >= 0 ? (nameToCopy >= 0 ? nameToCopy : left) : right;
            } else if (left >= 0 && right >= 0) {  // group.cpp:114-123
                if (b.reads[left].l_qname <= b.reads[right].l_qname) { gr.qname_donor[1] = left; name_slot = left; }
                else { gr.qname_donor[0] = right; name_slot = right; }
            } else {
                name_slot = left >= 0 ? left : right;
            }
            gr.umi_pair = name_slot >= 0 ? name_slot / 2 : -1;  // Pair::setLeft/setRight, pair.cpp:188-216
        }
        g.sync();
        for (int s = 0; s < 2; s++) {
            FsDesc fd = {0, 0, 0, 0, 0, SIDE_NONE, 0, 0, 0, 0, 0};
            if (gr.tmpl_read[s] >= 0) {
                const gcb_read_desc od = b.reads[gr.tmpl_read[s]];
                const int l_out = od.l_qseq;
                fd.c = c;
                const uint32_t *ocig = b.cigar + od.cigar_off;
                if (od.isize != 0 && gv.packed4) {  // group.cpp:362-367 + reference.cpp:33-71
                    const int contig = b.cluster_ref[c];
                    if (contig >= 0 && contig < gv.n_contigs) {
                        const int64_t span = (int64_t)get_ref_offset(ocig, od.n_cigar, ch[s].len - 1) + 1;
                        if ((int64_t)od.pos + span < gv.contig_len[contig]) {
                            fd.flags |= FS_REF_OK;
                            fd.ref_nib0 = 2 * gv.contig_off[contig] + od.pos;
                        }
                    }
                }
                if (od.n_cigar == 1) {
                    const int op = cig_op(ocig[0]);
                    if (query_consum(op) && ref_consum(op) && cig_len(ocig[0]) >= l_out) fd.flags |= FS_SIMPLE_CIGAR;
                }
                gr.out_off[s] = out_rel;  // cluster-relative; the vote kernel rebases it after the scan
                fd.mb = mb;
                fd.m = (uint16_t)m;
                if (ch[s].uniform) fd.flags |= FS_UNIFORM;
                fd.l_out = (uint16_t)l_out;
                fd.len = (uint16_t)ch[s].len;
                fd.tmpl_k = (uint16_t)ch[s].k;
                fd.mode = s == 0 && mode0 == SIDE_COPY ? SIDE_COPY : ws.side_mode[2 * (int64_t)slot + s];
                fd.out_rel = (uint32_t)out_rel;
                if (!ch[s].fits || m > 0xFFFF || l_out > 0x7FFF || out_rel > 0xFFFFFFFFll) fd.flags |= FS_NOFIT;
                out_rel += record_bytes(l_out);
            }
            if (lane == 0) ws.fs_desc[2 * (int64_t)slot + s] = fd;
        }
        if (lane == 0) r.groups[slot] = gr;
    }
    if (lane == 0) ws.cluster_out_bytes[c] = out_rel;
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan of cluster_out_bytes in two launches (block-local prefix, then the block totals).
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = SCAN_BLOCK / SCAN_THREADS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(Workspace ws, int32_t n_clusters) {
    GCB_GRID_DEP();
    __shared__ int64_t warp_tot[SCAN_THREADS / WARP];
    const int lane = lane_id(), warp = (int)(threadIdx.x >> 5);
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n_clusters ? ws.cluster_out_bytes[base + k] : 0;
        sum += v[k];
    }
    int64_t incl = sum;
    for (int off = 1; off < WARP; off <<= 1) {
        const int64_t t = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == WARP - 1) warp_tot[warp] = incl;
    __syncthreads();
    int64_t pre = incl - sum;
    for (int w = 0; w < warp; w++) pre += warp_tot[w];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n_clusters) ws.cluster_out_off[base + k] = pre;
        pre += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1) ws.scan_block[blockIdx.x] = pre;
}

// `carry_in`: bytes emitted by the views that precede this one in out_payload (NULL = none)
__global__ void __launch_bounds__(WARP) scan_blocks_kernel(Workspace ws, int32_t n_blocks, int64_t *out_bytes, int64_t out_capacity,
                                                           const int64_t *carry_in) {
    GCB_GRID_DEP();
    const int lane = lane_id();
    int64_t carry = carry_in ? *carry_in : 0;
    for (int base = 0; base < n_blocks; base += WARP) {
        const int i = base + lane;
        const int64_t v = i < n_blocks ? ws.scan_block[i] : 0;
        int64_t incl = v;
        for (int off = 1; off < WARP; off <<= 1) {
            const int64_t t = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += t;
        }
        if (i < n_blocks) ws.scan_block[i] = carry + incl - v;
        carry += __shfl_sync(FULL, incl, WARP - 1);
    }
    if (lane == 0) {
        ws.scan_block[n_blocks] = carry;
        *out_bytes = carry;
        if (carry > out_capacity) raise_error(ws.error_flag, GCB_ERR_CAPACITY);
    }
}

}  // namespace gcb
