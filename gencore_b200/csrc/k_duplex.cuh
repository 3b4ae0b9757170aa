// k_duplex.cuh — the tail of Cluster::clusterByUMI (cluster.cpp:102-188): duplex partner search,
// Cluster::duplexMerge / duplexMergeBam (cluster.cpp:190-244) on the consensus records the vote
// kernel wrote, and the SSCS / DCS / dropped verdict of every family.  Eight or sixteen lanes per cluster, a family
// per lane (duplex_kernel below); the partner search is a sequential stack walk in the reference and its order decides
// who pairs up.
#pragma once

#include "k_group_select.cuh"

namespace gcb {

constexpr int DUPLEX_THREADS = 128;

// cluster.cpp:200-244 on two consensus records in out_payload.  The byte-equality shortcut advances
// i by two from either parity, so after a handled mismatch the walk can stay on odd indices and skip
// the high nibbles of the following differing bytes; reproduced literally.
GCB_DEV int duplex_walk(uint8_t *rec1, int len1, uint8_t *rec2, int len2, int start) {  // the differing positions from `start` on
    int diff = 0;
    const int len = min(len1, len2);
    uint8_t *qual1 = rec1, *qual2 = rec2;
    uint8_t *seq1 = rec1 + GCB_ALIGN4(len1), *seq2 = rec2 + GCB_ALIGN4(len2);
    for (int i = start; i < len; i++) {
        const uint8_t a = seq1[i >> 1], c = seq2[i >> 1];
        if (a == c) {
            i++;
            continue;
        }
        const int b1 = (i & 1) ? (a & 0xF) : (a >> 4), b2 = (i & 1) ? (c & 0xF) : (c >> 4);
        if (base_letter(b1) != base_letter(b2)) {
            diff++;
            qual1[i] = 0;
            qual2[i] = 0;
            if (i & 1) {
                seq1[i >> 1] = (uint8_t)(a | 0x0F);
                seq2[i >> 1] = (uint8_t)(c | 0x0F);
            } else {
                seq1[i >> 1] = (uint8_t)(a | 0xF0);
                seq2[i >> 1] = (uint8_t)(c | 0xF0);
            }
        }
    }
    return diff;
}
GCB_DEV int duplex_merge_records(uint8_t *rec1, int len1, uint8_t *rec2, int len2) {
    return (len1 > len2 ? len1 - len2 : len2 - len1) + duplex_walk(rec1, len1, rec2, len2, 0);
}

// The same walk over copies of the two sequences in shared memory (the walk is a chain of dependent byte reads: from global
// memory every step costs an L2 round trip, from shared memory a few cycles).  s1 / s2: this thread's two staging rows of
// DUPLEX_STAGE_WORDS words.  The modifications go to global memory AND to the copies (the walk re-reads what it wrote).
constexpr int DUPLEX_STAGE_WORDS = 39;  // 156 bytes = 312 bases; odd, so that the threads' rows fall into different banks
GCB_DEV int duplex_merge_staged(uint8_t *rec1, int len1, uint8_t *rec2, int len2, uint32_t *s1, uint32_t *s2) {
    int diff = len1 > len2 ? len1 - len2 : len2 - len1;
    const int len = min(len1, len2);
    uint8_t *qual1 = rec1, *qual2 = rec2;
    uint8_t *seq1 = rec1 + GCB_ALIGN4(len1), *seq2 = rec2 + GCB_ALIGN4(len2);
    const int nw = (((len + 1) >> 1) + 3) >> 2;  // (records are 4-byte aligned and padded to whole words)
    for (int k = 0; k < nw; k++) {
        s1[k] = ((const uint32_t *)seq1)[k];
        s2[k] = ((const uint32_t *)seq2)[k];
    }
    uint8_t *b1 = (uint8_t *)s1, *b2 = (uint8_t *)s2;
    for (int i = 0; i < len; i++) {
        const uint8_t a = b1[i >> 1], c = b2[i >> 1];
        if (a == c) {
            i++;
            continue;
        }
        const int x1 = (i & 1) ? (a & 0xF) : (a >> 4), x2 = (i & 1) ? (c & 0xF) : (c >> 4);
        if (base_letter(x1) != base_letter(x2)) {
            diff++;
            qual1[i] = 0;
            qual2[i] = 0;
            const uint8_t m = (i & 1) ? 0x0F : 0xF0;
            b1[i >> 1] = (uint8_t)(a | m);
            b2[i >> 1] = (uint8_t)(c | m);
            seq1[i >> 1] = (uint8_t)(a | m);
            seq2[i >> 1] = (uint8_t)(c | m);
        }
    }
    return diff;
}

// The first byte at or after `from` in which the two sequences' first `nbytes` bytes differ, found by the group's lanes
// together, a word per lane and step (0x7FFFFFFF: none).  The records are 4-byte aligned and padded to whole words.
template <int GS>
GCB_DEV int first_diff_byte(const Grp<GS> &g, const uint8_t *seq1, const uint8_t *seq2, int nbytes, int from) {
    int k0 = 0x7FFFFFFF;
    for (int w = (from >> 2) + g.gl; 4 * w < nbytes && k0 == 0x7FFFFFFF; w += GS) {
        uint32_t x = ((const uint32_t *)seq1)[w] ^ ((const uint32_t *)seq2)[w];
        const int rem = nbytes - 4 * w, skip = from - 4 * w;
        if (rem < 4) x &= (1u << (8 * rem)) - 1u;
        if (skip > 0) x &= ~((1u << (8 * skip)) - 1u);  // (skip <= 3: only in the first word)
        if (x != 0u) k0 = 4 * w + ((__ffs((int)x) - 1) >> 3);
    }
    return g.min_of(k0);
}

// cluster.cpp:200-244 by a group of lanes: equal bytes only move the walk on by two positions, whatever its parity, so the
// lanes look for the next differing byte together and lane 0 runs the reference's loop literally inside that byte (one or
// two positions; its writes change this byte alone, which the walk then leaves).  Returns the differing positions.
template <int GS>
GCB_DEV int duplex_merge_group(const Grp<GS> &g, uint8_t *rec1, int len1, uint8_t *rec2, int len2) {
    const int len = min(len1, len2), nbytes = (len + 1) >> 1;
    uint8_t *qual1 = rec1, *qual2 = rec2;
    uint8_t *seq1 = rec1 + GCB_ALIGN4(len1), *seq2 = rec2 + GCB_ALIGN4(len2);
    int diff = 0, i = 0;
    for (;;) {
        const int kb = first_diff_byte(g, seq1, seq2, nbytes, i >> 1);
        if (kb == 0x7FFFFFFF) break;
        if (kb > (i >> 1)) i = 2 * kb + (i & 1);
        if (g.gl == 0) {
            for (; i < len && (i >> 1) == kb; i++) {
                const uint8_t a = seq1[i >> 1], c = seq2[i >> 1];
                if (a == c) {
                    i++;
                    continue;
                }
                const int b1 = (i & 1) ? (a & 0xF) : (a >> 4), b2 = (i & 1) ? (c & 0xF) : (c >> 4);
                if (base_letter(b1) != base_letter(b2)) {
                    diff++;
                    qual1[i] = 0;
                    qual2[i] = 0;
                    const uint8_t m = (i & 1) ? 0x0F : 0xF0;
                    seq1[i >> 1] = (uint8_t)(a | m);
                    seq2[i >> 1] = (uint8_t)(c | m);
                }
            }
        }
        i = __shfl_sync(g.mask, i, g.base);
        if (i >= len) break;
    }
    return __shfl_sync(g.mask, diff, g.base);
}

// The reference's loop over the stack for one cluster by two threads (side = 0 / 1): both walk the stack (same decisions),
// each merges one side of a strand pair's consensus records, thread 0 writes the verdicts.  For clusters with more
// families than a group has lanes.
GCB_DEV void duplex_cluster_pairwise(const BatchView &b, const ResultView &r, const Workspace &ws, const gcb_options &o, int c, int side, unsigned pairmask,
                                     uint32_t *stage1, uint32_t *stage2) {
    const int p0 = b.cluster_pair_off[c];
    const int G = r.cluster_n_groups[c];
    const int nw = b.umi_words;
    gcb_group_result *gr = r.groups + p0;
    // cluster.cpp:119-168: pop from the back, pair with the first family (in creation order) whose UMI is the swap
    int32_t *alive = ws.scratch + 2 * (int64_t)p0;  // G <= pairs of the cluster
    int nalive = G;
    if (side == 0)
        for (int g = 0; g < G; g++) alive[g] = g;
    __syncwarp(pairmask);
    while (nalive > 0) {
        const int g1 = alive[--nalive];
        gcb_group_result *r1 = gr + g1;
        const Umi u1 = r1->umi_pair >= 0 ? umi_load(b.umi + (int64_t)r1->umi_pair * nw, nw) : umi_load(b.umi, 0);
        bool found = false;
        for (int i = 0; i < nalive; i++) {
            const int g2 = alive[i];
            gcb_group_result *r2 = gr + g2;
            const Umi u2 = r2->umi_pair >= 0 ? umi_load(b.umi + (int64_t)r2->umi_pair * nw, nw) : umi_load(b.umi, 0);
            if (!umi_is_duplex(u1, u2)) continue;
            found = true;
            int diff = 0;  // Cluster::duplexMerge, cluster.cpp:190-198: this thread's side
            {
                const int s = side;
                const int t1 = r1->tmpl_read[s], t2 = r2->tmpl_read[s];
                if (t1 >= 0 && t2 >= 0) {
                    const int l1 = b.reads[t1].l_qseq, l2 = b.reads[t2].l_qseq;
                    if (!(r1->out_off[s] + record_bytes(l1) > r.out_capacity || r2->out_off[s] + record_bytes(l2) > r.out_capacity)) {
                        if (min(l1, l2) <= 8 * DUPLEX_STAGE_WORDS)
                            diff = duplex_merge_staged(r.out_payload + r1->out_off[s], l1, r.out_payload + r2->out_off[s], l2, stage1, stage2);
                        else
                            diff = duplex_merge_records(r.out_payload + r1->out_off[s], l1, r.out_payload + r2->out_off[s], l2);
                    }
                }
            }
            diff += __shfl_xor_sync(pairmask, diff, 1);
            if (side == 0) {
                r1->duplex_partner = g2;
                r1->duplex_diff = diff;
                r2->duplex_partner = g1;
                r2->duplex_diff = diff;
                r2->status = GCB_GROUP_DUPLEX_PARTNER;
                if (diff <= o.duplex_mismatch_threshold) {
                    if (r1->merge_reads + r2->merge_reads >= o.cluster_size_req) {
                        r1->status = GCB_GROUP_DCS;
                        r1->reverse_merge_reads = r2->merge_reads;  // Pair::setDuplex
                    } else {
                        r1->status = GCB_GROUP_DUPLEX_SMALL;
                    }
                } else {
                    r1->status = GCB_GROUP_DUPLEX_DIFF;
                }
                for (int k = i; k + 1 < nalive; k++) alive[k] = alive[k + 1];
            }
            __syncwarp(pairmask);
            nalive--;
            break;
        }
        if (!found && side == 0) r1->status = (!o.duplex_only && r1->merge_reads >= o.cluster_size_req) ? GCB_GROUP_SSCS : GCB_GROUP_DROPPED;
    }
}

// DUPLEX_GS lanes per cluster (eight or sixteen), lane i holds family i (its UMI, its sizes, where its consensus records lie).  The stack of
// cluster.cpp:119-168 keeps the families in creation order, so it is a bit mask: the popped family is its highest bit, the
// partner the lowest bit whose UMI is the swap — every lane tests its own family, one ballot finds the partner.  The merge
// (cluster.cpp:200-244) is a walk whose outcome depends on its own writes, but only from the first differing byte on: the
// lanes find that byte together (a word each), and only strand pairs that differ anywhere are walked, by one lane, from there.
// (Clusters with more families than lanes take the two-lane walk: rare, and slow — the host gives sixteen lanes to batches of
// larger clusters: 0.154 -> 0.056 ms on the cfg4 shape, where a few clusters in a thousand have nine or more families.)
template <int DUPLEX_GS>
__global__ void __launch_bounds__(DUPLEX_THREADS) duplex_kernel(BatchView b, ResultView r, Workspace ws, gcb_options o) {
    GCB_GRID_DEP();
    __shared__ uint32_t s_stage[DUPLEX_THREADS / DUPLEX_GS][2][2][DUPLEX_STAGE_WORDS];
    const Grp<DUPLEX_GS> g;
    const int lane = g.gl;
    const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) / DUPLEX_GS);
    if (c >= b.n_clusters || batch_is_malformed(ws.error_flag)) return;
    const int p0 = b.cluster_pair_off[c];
    const int G = r.cluster_n_groups[c];
    const int nw = b.umi_words;
    gcb_group_result *gr = r.groups + p0;

    if (!(ws.cluster_has_umi[c] && !o.disable_duplex)) {  // cluster.cpp:169-183
        for (int gi = lane; gi < G; gi += DUPLEX_GS)
            gr[gi].status = (!o.duplex_only && gr[gi].merge_reads >= o.cluster_size_req) ? GCB_GROUP_SSCS : GCB_GROUP_DROPPED;
        return;
    }
    if (G > DUPLEX_GS) {
        if (lane < 2) {
            uint32_t(*st)[2][DUPLEX_STAGE_WORDS] = s_stage[threadIdx.x / DUPLEX_GS];
            duplex_cluster_pairwise(b, r, ws, o, c, lane, 3u << g.base, st[lane][0], st[lane][1]);
        }
        return;
    }
    const bool fam = lane < G;
    int umi_pair = -1, merge_reads = 0, tlen[2] = {-1, -1};  // tlen: l_qseq of the family's template on either side, -1 = no record
    long long out_off[2] = {0, 0};
    if (fam) {
        const gcb_group_result *row = gr + lane;
        umi_pair = row->umi_pair;
        merge_reads = row->merge_reads;
        const int t0 = row->tmpl_read[0], t1 = row->tmpl_read[1];
        out_off[0] = row->out_off[0]; out_off[1] = row->out_off[1];
        if (t0 >= 0) tlen[0] = b.reads[t0].l_qseq;  // (loaded once per family, not once per merge: the loop below is a chain of dependent loads as it is)
        if (t1 >= 0) tlen[1] = b.reads[t1].l_qseq;
    }
    Umi canon, swapped;  // cluster.cpp:246-258 isDuplex(a, b) <=> canon(a) == swapped(b)
    const bool two = umi_strand_forms(umi_load(b.umi + (int64_t)(umi_pair >= 0 ? umi_pair : 0) * nw, umi_pair >= 0 ? nw : 0), canon, swapped);
    unsigned alive = (1u << G) - 1u;
    int status = GCB_GROUP_DROPPED, partner = -1, ddiff = 0, rev = 0;
    while (alive != 0u) {
        const int g1 = 31 - __clz((int)alive);  // cluster.cpp:119-121: the back of the stack
        alive &= ~(1u << g1);
        Umi c1;
#pragma unroll
        for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) c1.w[k] = __shfl_sync(g.mask, canon.w[k], g.base + g1);
        const bool two1 = __shfl_sync(g.mask, (int)two, g.base + g1) != 0;
        const bool match = ((alive >> lane) & 1u) != 0u && two && two1 && umi_equal(c1, swapped);
        const unsigned bal = g.ballot(match);
        const int mr1 = __shfl_sync(g.mask, merge_reads, g.base + g1);
        if (bal == 0u) {
            if (lane == g1) status = (!o.duplex_only && mr1 >= o.cluster_size_req) ? GCB_GROUP_SSCS : GCB_GROUP_DROPPED;
            continue;
        }
        const int g2 = __ffs((int)bal) - 1;  // the first family in creation order whose UMI is the swap
        alive &= ~(1u << g2);
        int diff = 0;  // Cluster::duplexMerge, cluster.cpp:190-198
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int l1 = __shfl_sync(g.mask, tlen[s], g.base + g1), l2 = __shfl_sync(g.mask, tlen[s], g.base + g2);
            const long long o1 = __shfl_sync(g.mask, out_off[s], g.base + g1), o2 = __shfl_sync(g.mask, out_off[s], g.base + g2);
            if (l1 < 0 || l2 < 0) continue;
            if (o1 + record_bytes(l1) > r.out_capacity || o2 + record_bytes(l2) > r.out_capacity) continue;
            uint8_t *rec1 = r.out_payload + o1, *rec2 = r.out_payload + o2;
            diff += l1 > l2 ? l1 - l2 : l2 - l1;
            diff += duplex_merge_group(g, rec1, l1, rec2, l2);
        }
        const int mr2 = __shfl_sync(g.mask, merge_reads, g.base + g2);
        if (lane == g1) {
            partner = g2;
            ddiff = diff;
            if (diff <= o.duplex_mismatch_threshold) {
                if (mr1 + mr2 >= o.cluster_size_req) {
                    status = GCB_GROUP_DCS;
                    rev = mr2;  // Pair::setDuplex
                } else {
                    status = GCB_GROUP_DUPLEX_SMALL;
                }
            } else {
                status = GCB_GROUP_DUPLEX_DIFF;
            }
        } else if (lane == g2) {
            partner = g1;
            ddiff = diff;
            status = GCB_GROUP_DUPLEX_PARTNER;
        }
    }
    if (fam) {
        gcb_group_result *row = gr + lane;
        row->status = status;
        if (partner >= 0) {
            row->duplex_partner = partner;
            row->duplex_diff = ddiff;
        }
        if (status == GCB_GROUP_DCS) row->reverse_merge_reads = rev;
    }
}

}  // namespace gcb
