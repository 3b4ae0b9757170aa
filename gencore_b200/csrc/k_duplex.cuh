// k_duplex.cuh — the tail of Cluster::clusterByUMI (cluster.cpp:102-188): duplex partner search,
// Cluster::duplexMerge / duplexMergeBam (cluster.cpp:190-244) on the consensus records the vote
// kernel wrote, and the SSCS / DCS / dropped verdict of every family.  Two threads per cluster: the
// partner search is a sequential stack walk in the reference and its order decides who pairs up.
#pragma once

#include "k_group_select.cuh"

namespace gcb {

constexpr int DUPLEX_THREADS = 128;

// cluster.cpp:200-244 on two consensus records in out_payload.  The byte-equality shortcut advances
// i by two from either parity, so after a handled mismatch the walk can stay on odd indices and skip
// the high nibbles of the following differing bytes; reproduced literally.
GCB_DEV int duplex_merge_records(uint8_t *rec1, int len1, uint8_t *rec2, int len2) {
    int diff = len1 > len2 ? len1 - len2 : len2 - len1;
    const int len = min(len1, len2);
    uint8_t *qual1 = rec1, *qual2 = rec2;
    uint8_t *seq1 = rec1 + GCB_ALIGN4(len1), *seq2 = rec2 + GCB_ALIGN4(len2);
    for (int i = 0; i < len; i++) {
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

// Two threads per cluster: both walk the stack (same decisions), each merges one side of a strand pair's consensus records
// (the walk over a record is sequential, the two sides are independent), thread 0 writes the verdicts.
__global__ void __launch_bounds__(DUPLEX_THREADS) duplex_kernel(BatchView b, ResultView r, Workspace ws, gcb_options o) {
    GCB_GRID_DEP();
    __shared__ uint32_t s_stage[DUPLEX_THREADS][2][DUPLEX_STAGE_WORDS];
    const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int c = t >> 1, side = t & 1;
    if (c >= b.n_clusters || batch_is_malformed(ws.error_flag)) return;
    const unsigned pairmask = 3u << (lane_id() & ~1);  // this cluster's two lanes
    const int p0 = b.cluster_pair_off[c];
    const int G = r.cluster_n_groups[c];
    const int nw = b.umi_words;
    gcb_group_result *gr = r.groups + p0;

    if (!(ws.cluster_has_umi[c] && !o.disable_duplex)) {  // cluster.cpp:169-183
        if (side == 0)
            for (int g = 0; g < G; g++)
                gr[g].status = (!o.duplex_only && gr[g].merge_reads >= o.cluster_size_req) ? GCB_GROUP_SSCS : GCB_GROUP_DROPPED;
        return;
    }
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
                            diff = duplex_merge_staged(r.out_payload + r1->out_off[s], l1, r.out_payload + r2->out_off[s], l2, s_stage[threadIdx.x][0],
                                                       s_stage[threadIdx.x][1]);
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

}  // namespace gcb
