// k_cluster_stats.cuh — the Stats side effects of Cluster::clusterByUMI (cluster.cpp:102,136,143,157,161,172,176,184-186 on
// stats.cpp:122-141: addCluster / addMolecule / addSSCS / addDCS) as one pass over the result rows: one THREAD per cluster,
// counters summed per warp and added to the context's accumulator with one atomic per counter and warp.
#pragma once

#include "k_group_select.cuh"

namespace gcb {

constexpr int STATS_THREADS = 256;

// indices into gcb_cluster_stats seen as int64[10 + GCB_MAX_SUPPORTING_READS]
enum { ST_PRE_CLUSTER = 0, ST_PRE_MULTI, ST_PRE_MOLECULE, ST_PRE_SE, ST_PRE_PE, ST_PRE_UNCOUNTED, ST_POST_CLUSTER, ST_POST_MULTI, ST_POST_SSCS,
       ST_POST_DCS, ST_HIST0 };

GCB_DEV void stats_add(int *cta_counters, int k, int v) {  // (the CTA's share: shared memory, flushed once per CTA)
    const int s = __reduce_add_sync(FULL, v);
    if (lane_id() == 0 && s != 0) atomicAdd(cta_counters + k, s);
}

__global__ void __launch_bounds__(STATS_THREADS) cluster_stats_kernel(BatchView b, ResultView r, Workspace ws, unsigned long long *acc) {
    GCB_GRID_DEP();
    __shared__ int s_cnt[ST_HIST0 + GCB_MAX_SUPPORTING_READS];  // the CTA's share of the counters and of Stats::mSupportingHistgram (a few
                                                                // hot words: thousands of global atomics on them would serialise)
    int *s_hist = s_cnt + ST_HIST0;
    for (int i = (int)threadIdx.x; i < ST_HIST0 + GCB_MAX_SUPPORTING_READS; i += (int)blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    if (batch_is_malformed(ws.error_flag)) return;
    const int n_round = (b.n_clusters + WARP - 1) / WARP * WARP;  // whole warps take part in the reductions
    for (int c = (int)(blockIdx.x * blockDim.x + threadIdx.x); c < n_round; c += (int)(gridDim.x * blockDim.x)) {
        int pre_multi = 0, mol = 0, pe = 0, uncounted = 0, kept = 0, sscs = 0, dcs = 0;
        const bool have = c < b.n_clusters;
        if (have) {
            const int ng = r.cluster_n_groups[c];
            const gcb_group_result *g0 = r.groups + b.cluster_pair_off[c];
            pre_multi = ng > 1;  // cluster.cpp:102
            for (int g = 0; g < ng; g++) {
                const gcb_group_result *gr = g0 + g;
                const int st = gr->status;
                if (st != GCB_GROUP_DUPLEX_PARTNER) {  // every group is one addMolecule, except a consumed duplex partner
                    mol++;
                    int supporting = gr->merge_reads;
                    if (gr->duplex_partner >= 0) supporting += g0[gr->duplex_partner].merge_reads;  // cluster.cpp:136,157
                    if (gr->tmpl_read[0] >= 0 && gr->tmpl_read[1] >= 0) pe++;
                    if ((unsigned)supporting < (unsigned)GCB_MAX_SUPPORTING_READS) atomicAdd(&s_hist[supporting], 1);  // stats.cpp:124-127
                    else uncounted++;
                }
                if (st == GCB_GROUP_SSCS) sscs++;
                if (st == GCB_GROUP_DCS) dcs++;
            }
            kept = sscs + dcs;
        }
        stats_add(s_cnt, ST_PRE_CLUSTER, have ? 1 : 0);
        stats_add(s_cnt, ST_PRE_MULTI, pre_multi);
        stats_add(s_cnt, ST_PRE_MOLECULE, mol);
        stats_add(s_cnt, ST_PRE_PE, pe);
        stats_add(s_cnt, ST_PRE_SE, mol - pe);
        stats_add(s_cnt, ST_PRE_UNCOUNTED, uncounted);
        stats_add(s_cnt, ST_POST_CLUSTER, kept > 0 ? 1 : 0);  // cluster.cpp:184-186
        stats_add(s_cnt, ST_POST_MULTI, kept > 1 ? 1 : 0);
        stats_add(s_cnt, ST_POST_SSCS, sscs);
        stats_add(s_cnt, ST_POST_DCS, dcs);
    }
    __syncthreads();
    for (int i = (int)threadIdx.x; i < ST_HIST0 + GCB_MAX_SUPPORTING_READS; i += (int)blockDim.x)
        if (s_cnt[i] != 0) atomicAdd(acc + i, (unsigned long long)s_cnt[i]);
}

}  // namespace gcb
