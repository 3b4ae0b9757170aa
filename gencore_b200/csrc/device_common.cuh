// device_common.cuh — shared device-side vocabulary of the consensus engine (sm_100a only).
// Reference semantics are cited as file:line under /root/reference/src.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gencore_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gencore_b200 is written for sm_100a (B200) only"
#endif

namespace gcb {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

// ---- workspace that lives in the context and carries state between the four kernels
struct PairOverlap {   // Pair::computeScore's overlap window (pair.cpp:103-119), one per pair
    int16_t left_start;   // first overlapped index in the left read
    int16_t right_start;  // first overlapped index in the right read
    int16_t cmp_len;      // overlapped length (may be <= 0)
    int16_t valid;        // 1: both mates present and both have an M block (scores are quality-derived)
};

struct Workspace {
    int32_t *members;         // [n_pairs] pair indices of each cluster, grouped by UMI family (stable)
    int32_t *group_start;     // [n_pairs] slot-indexed: offset (within the batch) of the family's first member
    int32_t *umi_count;       // [n_pairs] multiplicity of each pair's UMI inside its cluster
    int32_t *right_ref_pos;   // [2*n_pairs] BamUtil::getRightRefPos per read
    uint8_t *vote_flags;      // [2*n_pairs] VOTE_* per read
    uint8_t *side_mode;       // [2*n_pairs] slot*2+side: leftReadMode used for that side's vote
    uint8_t *cluster_has_umi; // [n_clusters]
    PairOverlap *overlap;     // [n_pairs]
    int64_t *cluster_out_bytes; // [n_clusters] bytes of consensus records the cluster emits
    int64_t *cluster_out_off;   // [n_clusters] exclusive scan of the above
    int64_t *scan_tiles;        // scratch of the scan
    int32_t *error_flag;        // [1] sticky gcb_status raised by a kernel
};

constexpr uint8_t VOTE_PARTICIPATES = 1;  // read is in makeConsensus' `reads` (group.cpp:287-313)
constexpr uint8_t VOTE_LENDIFF0 = 2;      // lenDiff forced to 0 (group.cpp:344-347)

struct GenomeView {
    const uint8_t *packed4;
    const int64_t *contig_off;
    const int64_t *contig_len;
    int32_t n_contigs;
};

// ---- CIGAR helpers ------------------------------------------------------------------------
__device__ __forceinline__ int cig_op(uint32_t c) { return (int)(c & 0xF); }
__device__ __forceinline__ int cig_len(uint32_t c) { return (int)(c >> 4); }
// bamutil.cpp:290-291 as bit masks over the op code (ops >= 10 consume nothing)
__device__ __forceinline__ int query_consum(int op) { return (0x193 >> op) & 1; }  // M I S = X
__device__ __forceinline__ int ref_consum(int op) { return (0x18D >> op) & 1; }    // M D N = X
constexpr int OP_MATCH = 0, OP_INS = 1, OP_SOFT_CLIP = 4, OP_HARD_CLIP = 5;

// BamUtil::getRefOffset, bamutil.cpp:293-314
__device__ __forceinline__ int get_ref_offset(const uint32_t *__restrict__ cig, int n, int bampos) {
    int ref = 0, query = 0;
    for (int i = 0; i < n; i++) {
        uint32_t v = __ldg(cig + i);
        int op = cig_op(v), len = cig_len(v);
        query += len * query_consum(op);
        ref += len * ref_consum(op);
        if (query > bampos) {
            if (op == OP_INS || op == OP_SOFT_CLIP) return -1;
            return ref - ref_consum(op) * (query - bampos);
        }
    }
    return -1;
}

// BamUtil::getMOffsetAndLen, bamutil.cpp:316-336
__device__ __forceinline__ void get_m_offset_and_len(const uint32_t *__restrict__ cig, int n, int &off, int &len) {
    int query = 0;
    for (int i = 0; i < n; i++) {
        uint32_t v = __ldg(cig + i);
        int op = cig_op(v);
        if (op == OP_MATCH) { off = query; len = cig_len(v); return; }
        query += cig_len(v) * query_consum(op);
    }
    off = 0;
    len = 0;
}

// bam_cigar2rlen (used by BamUtil::getRightRefPos, bamutil.cpp:379-383)
__device__ __forceinline__ int cigar_ref_len(const uint32_t *__restrict__ cig, int n) {
    int l = 0;
    for (int i = 0; i < n; i++) {
        uint32_t v = __ldg(cig + i);
        l += cig_len(v) * ref_consum(cig_op(v));
    }
    return l;
}

// BamUtil::isPartOf, bamutil.cpp:204-255
__device__ __forceinline__ bool is_part_of(const uint32_t *__restrict__ cp, int np, const uint32_t *__restrict__ cw, int nw,
                                           bool is_left) {
    if (nw < np) return false;
    for (int i = 0; i < np; i++) {
        uint32_t vp = __ldg(is_left ? cp + i : cp + (np - i - 1));
        uint32_t vw = __ldg(is_left ? cw + i : cw + (nw - i - 1));
        if (cig_op(vp) != cig_op(vw)) return false;
        if (cig_len(vp) > cig_len(vw)) return false;
        if (cig_len(vp) < cig_len(vw)) {
            if (i != np - 1) {
                if (i != np - 2) return false;
                uint32_t vn = __ldg(is_left ? cp + i + 1 : cp + (np - i - 2));
                if (cig_op(vn) != OP_HARD_CLIP) return false;
            }
        }
    }
    return true;
}

// ---- UMI code helpers (see the header's encoding conventions) -------------------------------
// number of differing 4-bit fields == Cluster::umiDiff (cluster.cpp:41-53)
__device__ __forceinline__ int umi_diff_word(uint64_t a, uint64_t b) {
    uint64_t x = a ^ b;
    x |= x >> 1;
    x |= x >> 2;
    return __popcll(x & 0x1111111111111111ull);
}

template <int W>
struct Umi {
    uint64_t w[W];
    __device__ __forceinline__ void load(const uint64_t *__restrict__ p) {
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = __ldg(p + k);
    }
    __device__ __forceinline__ int diff(const Umi &o) const {
        int d = 0;
#pragma unroll
        for (int k = 0; k < W; k++) d += umi_diff_word(w[k], o.w[k]);
        return d;
    }
    __device__ __forceinline__ bool equal(const Umi &o) const {
        bool e = true;
#pragma unroll
        for (int k = 0; k < W; k++) e &= (w[k] == o.w[k]);
        return e;
    }
    // std::string operator< on the decoded UMIs
    __device__ __forceinline__ bool less(const Umi &o) const {
#pragma unroll
        for (int k = 0; k < W; k++) {
            if (w[k] != o.w[k]) return w[k] < o.w[k];
        }
        return false;
    }
    __device__ __forceinline__ bool empty() const { return (w[0] >> 60) == 0; }
};

// nibble k (0 = first character) of a UMI code held in global memory
__device__ __forceinline__ int umi_field(const uint64_t *__restrict__ u, int k) {
    return (int)((__ldg(u + (k >> 4)) >> (60 - 4 * (k & 15))) & 0xF);
}

// ---- sequence helpers ------------------------------------------------------------------------
__device__ __forceinline__ int base_at(const uint8_t *seq, int i) {  // bam_get_seq nibble order
    uint8_t b = seq[i >> 1];
    return (i & 1) ? (b & 0xF) : (b >> 4);
}

// Pair::qual2score, pair.cpp:77-86
__device__ __forceinline__ int qual2score(const gcb_options &o, int q) {
    if (o.high_quality <= q) return o.score_high;
    if (o.moderate_quality <= q) return o.score_moderate;
    if (o.low_quality <= q) return o.score_low;
    return o.score_bad;
}

// record size of a read of l bases in a payload
__device__ __forceinline__ int64_t record_bytes(int l) { return GCB_ALIGN4(l) + GCB_ALIGN4((l + 1) >> 1); }

__device__ __forceinline__ void raise_error(int32_t *flag, int code) { atomicCAS(flag, 0, code); }

}  // namespace gcb
