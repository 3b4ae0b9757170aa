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

// One byte of that walk, entered at index parity p (0: at its high nibble, index 2B; 1: at its low nibble, index 2B + 1).
// Returns the parity at which the next byte is entered.  `a` / `c` are the two records' bytes and come back masked;
// with `apply` the mismatches are counted and the qualities zeroed.
GCB_DEV int duplex_byte(uint32_t &a, uint32_t &c, int B, int p, int len, bool apply, int &diff, uint8_t *q1, uint8_t *q2) {
    if (p == 0) {
        if (2 * B >= len || a == c) return 0;
        if (base_letter((int)(a >> 4)) != base_letter((int)(c >> 4))) {
            if (apply) {
                diff++;
                q1[2 * B] = 0;
                q2[2 * B] = 0;
            }
            a |= 0xF0u;
            c |= 0xF0u;
        }
    }
    if (2 * B + 1 >= len || a == c) return 1;
    if (base_letter((int)(a & 0xFu)) != base_letter((int)(c & 0xFu))) {
        if (apply) {
            diff++;
            q1[2 * B + 1] = 0;
            q2[2 * B + 1] = 0;
        }
        a |= 0x0Fu;
        c |= 0x0Fu;
    }
    return 0;
}

// The walk of duplex_merge_records by NL lanes (`seg` = this lane, `lanes` = their mask): every lane takes a run of
// bytes, finds where the walk leaves its run for both ways of entering it, the entry parities follow from a ballot, and
// every lane then walks its run for real.  Returns this lane's share of the mismatches.
template <int NL>
GCB_DEV int duplex_merge_records_par(uint8_t *rec1, int len1, uint8_t *rec2, int len2, int seg, unsigned lanes, int first_lane) {
    const int len = min(len1, len2);
    uint8_t *seq1 = rec1 + GCB_ALIGN4(len1), *seq2 = rec2 + GCB_ALIGN4(len2);
    const int nbytes = (len + 1) >> 1, per = (nbytes + NL - 1) / NL;
    const int b0 = min(seg * per, nbytes), b1 = min(b0 + per, nbytes);
    int exit_par[2], unused = 0;
#pragma unroll
    for (int p0 = 0; p0 < 2; p0++) {
        int p = p0;
        for (int B = b0; B < b1; B++) {
            uint32_t a = seq1[B], c = seq2[B];
            p = duplex_byte(a, c, B, p, len, false, unused, rec1, rec2);
        }
        exit_par[p0] = p;
    }
    const unsigned e0 = (__ballot_sync(lanes, exit_par[0] != 0) >> first_lane) & ((1u << NL) - 1u);
    const unsigned e1 = (__ballot_sync(lanes, exit_par[1] != 0) >> first_lane) & ((1u << NL) - 1u);
    int p = 0;
    for (int k = 0; k < seg; k++) p = (int)(((p ? e1 : e0) >> k) & 1u);
    int diff = 0;
    for (int B = b0; B < b1; B++) {
        const uint32_t a0 = seq1[B], c0 = seq2[B];
        uint32_t a = a0, c = c0;
        p = duplex_byte(a, c, B, p, len, true, diff, rec1, rec2);
        if (a != a0) seq1[B] = (uint8_t)a;
        if (c != c0) seq2[B] = (uint8_t)c;
    }
    return diff;
}

// LPC lanes per cluster: all walk the stack (same decisions), half of them merge each side of a strand pair's consensus
// records, the first writes the verdicts.  LPC = 2 for batches that cannot hold duplex UMIs (one word), 16 otherwise.
template <int LPC>
__global__ void __launch_bounds__(DUPLEX_THREADS) duplex_kernel(BatchView b, ResultView r, Workspace ws, gcb_options o) {
    constexpr int NL = LPC / 2;  // lanes per side
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(t / LPC), sub = (int)(t % LPC), side = sub / NL, seg = sub % NL;
    if (c >= b.n_clusters) return;
    const int first = lane_id() & ~(LPC - 1);
    const unsigned cmask = (LPC == 32 ? 0xffffffffu : ((1u << LPC) - 1u)) << first;  // this cluster's lanes
    const int p0 = b.cluster_pair_off[c];
    const int G = r.cluster_n_groups[c];
    const int nw = b.umi_words;
    gcb_group_result *gr = r.groups + p0;

    if (!(ws.cluster_has_umi[c] && !o.disable_duplex)) {  // cluster.cpp:169-183
        for (int g = sub; g < G; g += LPC)
            gr[g].status = (!o.duplex_only && gr[g].merge_reads >= o.cluster_size_req) ? GCB_GROUP_SSCS : GCB_GROUP_DROPPED;
        return;
    }
    // cluster.cpp:119-168: pop from the back, pair with the first family (in creation order) whose UMI is the swap
    int32_t *alive = ws.scratch + 2 * (int64_t)p0;  // G <= pairs of the cluster
    int nalive = G;
    for (int g = sub; g < G; g += LPC) alive[g] = g;
    __syncwarp(cmask);
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
            int diff = 0;  // Cluster::duplexMerge, cluster.cpp:190-198: this lane's side
            {
                const int s = side;
                const int t1 = r1->tmpl_read[s], t2 = r2->tmpl_read[s];
                if (t1 >= 0 && t2 >= 0) {
                    const int l1 = b.reads[t1].l_qseq, l2 = b.reads[t2].l_qseq;
                    if (!(r1->out_off[s] + record_bytes(l1) > r.out_capacity || r2->out_off[s] + record_bytes(l2) > r.out_capacity)) {
                        uint8_t *q1 = r.out_payload + r1->out_off[s], *q2 = r.out_payload + r2->out_off[s];
                        if (NL == 1) {
                            diff = duplex_merge_records(q1, l1, q2, l2);
                        } else {
                            const unsigned lanes = ((1u << NL) - 1u) << (first + side * NL);
                            diff = duplex_merge_records_par<NL>(q1, l1, q2, l2, seg, lanes, first + side * NL);
                            if (seg == 0) diff += l1 > l2 ? l1 - l2 : l2 - l1;
                        }
                    }
                }
            }
            diff = __reduce_add_sync(cmask, diff);
            if (sub == 0) {
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
            __syncwarp(cmask);
            nalive--;
            break;
        }
        if (!found && sub == 0) r1->status = (!o.duplex_only && r1->merge_reads >= o.cluster_size_req) ? GCB_GROUP_SSCS : GCB_GROUP_DROPPED;
    }
}

}  // namespace gcb
