// device_common.cuh — shared device-side vocabulary of the consensus engine (sm_100a only).
// Reference semantics are cited as file:line under /root/reference/src.
#pragma once

#include <stdint.h>

#include "gencore_b200.h"
#include "simt.h"

namespace gcb {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

// ---- workspace that lives in the context and carries state between the kernels of one batch
struct PairOverlap {   // Pair::computeScore's overlap window (pair.cpp:103-119), one per pair
    int32_t left_start;   // first overlapped index in the left read
    int32_t right_start;  // first overlapped index in the right read
    int32_t cmp_len;      // overlapped length (may be <= 0)
    int32_t valid;        // 1: both mates present and both have an M block (scores depend on quality)
};

// One read as the tiled vote kernel sees it (16 bytes, staged into shared memory by a bulk copy).
// Offsets are in 4-byte units relative to the cluster's slab.  own_l == 0: the read does not vote.
struct __align__(16) VoteRead {
    uint16_t own_off4;   // record of this read
    uint16_t mate_off4;  // record of its mate (valid with VR_OVERLAP)
    int16_t own_l;       // l_qseq of this read, 0 = no vote
    int16_t mate_l;
    int16_t shift;       // readpos = column + shift (group.cpp:377-379)
    int16_t ov_own;      // overlap window of Pair::computeScore (pair.cpp:108-119) clipped to this read:
    int16_t ov_mate;     //   own index ov_own + k pairs with mate index ov_mate + k, 0 <= k < ov_len
    int16_t ov_len;      //   VR_NO_OVERLAP_INFO: pair without mate or M block, the moderate score everywhere (pair.cpp:92,99)
};
constexpr int VR_NO_OVERLAP_INFO = -1;

// One (family, side) for the tiled vote kernel (32 bytes, index 2*slot+side); self-contained, so that the
// vote kernel needs no cluster lookup before it can place the family side in its tile.
struct __align__(16) FsDesc {
    int32_t mb;        // members index of the family's first pair
    uint16_t m;        // pairs in the family
    uint16_t l_out;    // template l_qseq
    uint16_t len;      // columns that are voted (group.cpp:354-360)
    uint16_t tmpl_k;   // template = k-th member of the family
    uint8_t mode;      // SIDE_*
    uint8_t flags;     // FS_*
    uint16_t reserved;
    uint32_t out_rel;  // consensus record offset relative to the cluster's output
    int32_t c;         // cluster
    int64_t ref_nib0;  // FS_REF_OK: nibble index in the packed genome of the template's core.pos
};
constexpr uint8_t FS_NOFIT = 1;         // some field does not fit its 16 bits: the tile goes to the generic kernel
constexpr uint8_t FS_REF_OK = 2;        // group.cpp:362-367 + reference.cpp:33-71: the vote may consult the reference
constexpr uint8_t FS_SIMPLE_CIGAR = 4;  // template CIGAR is one M/=/X op covering the read: BamUtil::getRefOffset(i) == i
constexpr uint8_t FS_UNIFORM = 8;       // every voter has the template's geometry: same length, no shift, same overlap window, same record-to-mate distance
constexpr uint8_t FS_SIDE1 = 0x80;      // (set by the vote kernel) this is the right-hand side of the pairs
constexpr uint16_t VR_NO_VOTE = 0xFFFFu;  // VoteRead.own_off4 of a read that does not vote

struct __align__(16) TileDir {
    int32_t c0;      // first cluster of the tile
    int32_t p0;      // its first pair
    int64_t slab0;   // payload offset of its slab
};

struct Workspace {
    int32_t *members;           // [n_pairs] pair indices, cluster by cluster, families contiguous, map order inside
    int32_t *group_off;         // [n_pairs] slot-indexed: index into members of the family's first pair
    int32_t *scratch;           // [2*n_pairs] per read slot: containedBy counts (group.cpp:196-233) / duplex stack
    int32_t *right_ref_pos;     // [2*n_pairs] BamUtil::getRightRefPos per read slot
    uint8_t *vote_flags;        // [2*n_pairs] VOTE_* per read slot
    uint8_t *side_mode;         // [2*n_pairs] index slot*2+side: SIDE_*
    uint8_t *cluster_has_umi;   // [n_clusters] cluster.cpp:57-65 hasUMI
    PairOverlap *overlap;       // [n_pairs]
    int64_t *slab_off;          // [n_clusters+1] payload byte offset of each cluster's slab
    int64_t *cluster_out_bytes; // [n_clusters] bytes of consensus records the cluster emits
    int64_t *cluster_out_off;   // [n_clusters] exclusive prefix of the above inside its scan block
    int64_t *scan_block;        // [n_scan_blocks+1] exclusive prefix over scan blocks; last = total
    int32_t *error_flag;        // [1] sticky gcb_status raised by a kernel
    VoteRead *vote_reads;       // [2*n_pairs] what the tiled vote kernel needs of every read, family-side runs:
                                //   the family whose pairs are members[mb..mb+m) owns entries [2*mb, 2*mb+m) for side 0
                                //   and [2*mb+m, 2*mb+2m) for side 1, in members order
    FsDesc *fs_desc;            // [2*n_pairs] index 2*slot+side: one family-side's vote parameters (mode NONE = nothing)
    TileDir *tile_dir;          // [n_tiles+1] first cluster / pair / slab byte of every vote tile
    int32_t *generic_tiles;     // [n_tiles] tiles the tiled vote kernel handed to the generic kernel
    int32_t *generic_count;     // [1]
};

constexpr uint8_t VOTE_PARTICIPATES = 1;  // read is in makeConsensus' `reads` (group.cpp:287-313)
constexpr uint8_t VOTE_LENDIFF0 = 2;      // lenDiff forced to 0 (group.cpp:344-347)

constexpr uint8_t SIDE_NONE = 0;   // consensusMergeBam returned NULL for this side
constexpr uint8_t SIDE_LEFT = 1;   // vote with left-aligned column indexing (leftReadMode)
constexpr uint8_t SIDE_RIGHT = 2;  // vote with right-aligned column indexing
constexpr uint8_t SIDE_COPY = 3;   // single pair without mRight: record passes through (group.cpp:73-77)

constexpr int SCAN_BLOCK = 2048;   // clusters per block of the output-offset scan

struct GenomeView {
    const uint8_t *packed4;
    const int64_t *contig_off;
    const int64_t *contig_len;
    int32_t n_contigs;
    int64_t packed_bytes;
};

// ---- CIGAR helpers ------------------------------------------------------------------------
GCB_HD int cig_op(uint32_t c) { return (int)(c & 0xF); }
GCB_HD int cig_len(uint32_t c) { return (int)(c >> 4); }
// bamutil.cpp:290-291 as bit masks over the op code (ops >= 10 consume nothing)
GCB_HD int query_consum(int op) { return (0x193 >> op) & 1; }  // M I S = X
GCB_HD int ref_consum(int op) { return (0x18D >> op) & 1; }    // M D N = X
constexpr int OP_MATCH = 0, OP_INS = 1, OP_SOFT_CLIP = 4, OP_HARD_CLIP = 5;

// BamUtil::getRefOffset, bamutil.cpp:293-314
GCB_HD int get_ref_offset(const uint32_t *cig, int n, int bampos) {
    int ref = 0, query = 0;
    for (int i = 0; i < n; i++) {
        uint32_t v = cig[i];
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
GCB_HD void get_m_offset_and_len(const uint32_t *cig, int n, int &off, int &len) {
    int query = 0;
    for (int i = 0; i < n; i++) {
        uint32_t v = cig[i];
        int op = cig_op(v);
        if (op == OP_MATCH) { off = query; len = cig_len(v); return; }
        query += cig_len(v) * query_consum(op);
    }
    off = 0;
    len = 0;
}

// bam_cigar2rlen (used by BamUtil::getRightRefPos, bamutil.cpp:379-383)
GCB_HD int cigar_ref_len(const uint32_t *cig, int n) {
    int l = 0;
    for (int i = 0; i < n; i++) l += cig_len(cig[i]) * ref_consum(cig_op(cig[i]));
    return l;
}

// BamUtil::isPartOf, bamutil.cpp:204-255
GCB_HD bool is_part_of(const uint32_t *cp, int np, const uint32_t *cw, int nw, bool is_left) {
    if (nw < np) return false;
    for (int i = 0; i < np; i++) {
        uint32_t vp = is_left ? cp[i] : cp[np - i - 1];
        uint32_t vw = is_left ? cw[i] : cw[nw - i - 1];
        if (cig_op(vp) != cig_op(vw)) return false;
        if (cig_len(vp) > cig_len(vw)) return false;
        if (cig_len(vp) < cig_len(vw)) {
            if (i != np - 1) {
                if (i != np - 2) return false;
                uint32_t vn = is_left ? cp[i + 1] : cp[np - i - 2];
                if (cig_op(vn) != OP_HARD_CLIP) return false;
            }
        }
    }
    return true;
}

// BamUtil::getCigar string equality (bamutil.cpp:191-202): op CHARACTER + length; bam_cigar_opchr maps
// every op code >= 10 to '?', so those compare equal
GCB_HD bool same_cigar_string(const uint32_t *a, int na, const uint32_t *b, int nb) {
    if (na != nb) return false;
    for (int k = 0; k < na; k++) {
        int oa = cig_op(a[k]), ob = cig_op(b[k]);
        if (oa >= 10) oa = 10;
        if (ob >= 10) ob = 10;
        if (oa != ob || cig_len(a[k]) != cig_len(b[k])) return false;
    }
    return true;
}

// ---- UMI code helpers (see the header's encoding conventions) -------------------------------
// number of differing 4-bit fields == Cluster::umiDiff (cluster.cpp:41-53)
GCB_HD int nibble_diff64(uint64_t a, uint64_t b) {
    uint64_t x = a ^ b;
    x |= x >> 1;
    x |= x >> 2;
    x &= 0x1111111111111111ull;
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

struct Umi {
    uint64_t w[GCB_MAX_UMI_WORDS];
};
GCB_HD Umi umi_load(const uint64_t *p, int nw) {
    Umi u;
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) u.w[k] = k < nw ? p[k] : 0ull;
    return u;
}
GCB_HD int umi_diff(const Umi &a, const Umi &b) {
    int d = 0;
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) d += nibble_diff64(a.w[k], b.w[k]);
    return d;
}
GCB_HD bool umi_equal(const Umi &a, const Umi &b) {
    bool e = true;
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) e = e && (a.w[k] == b.w[k]);
    return e;
}
// std::string operator< on the decoded UMIs (MSB-first fields: plain unsigned word order)
GCB_HD bool umi_less(const Umi &a, const Umi &b) {
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++)
        if (a.w[k] != b.w[k]) return a.w[k] < b.w[k];
    return false;
}
GCB_HD int umi_field(const Umi &u, int k) { return (int)((u.w[k >> 4] >> (60 - 4 * (k & 15))) & 0xF); }
constexpr int UMI_UNDERSCORE = 5;

// Cluster::isDuplex (cluster.cpp:246-258) on the 4-bit code.  util.h:59-88 split(): leading '_' are
// skipped, then every '_' ends a part and the text after the last '_' is always a part (possibly
// empty); both UMIs must have exactly two parts, swapped.
struct UmiParts { int n, b0, e0, b1, e1; };
GCB_HD UmiParts umi_split(const Umi &u) {
    UmiParts p = {0, 0, 0, 0, 0};
    int len = 0;
    while (len < 16 * GCB_MAX_UMI_WORDS && umi_field(u, len) != 0) len++;
    if (len == 0) return p;
    int pos = 0;
    while (pos < len && umi_field(u, pos) == UMI_UNDERSCORE) pos++;
    if (pos >= len) return p;
    for (;;) {
        int sep = -1;
        for (int k = pos; k < len; k++)
            if (umi_field(u, k) == UMI_UNDERSCORE) { sep = k; break; }
        int end = sep >= 0 ? sep : len;
        if (p.n == 0) { p.b0 = pos; p.e0 = end; }
        else if (p.n == 1) { p.b1 = pos; p.e1 = end; }
        p.n++;
        if (sep < 0) break;
        pos = sep + 1;
    }
    return p;
}
// true when the code holds a '_' field: without one util.h:59-88 split() yields a single part and isDuplex is false
// (cluster.cpp:246-258), so the pairing loop can skip the split of every UMI of a single-strand library
GCB_HD bool umi_has_separator(const Umi &u) {
    bool any = false;
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) {
        const uint64_t y = u.w[k] ^ 0x5555555555555555ull;  // a zero nibble of y is a '_' field
        any = any || (((y - 0x1111111111111111ull) & ~y & 0x8888888888888888ull) != 0ull);
    }
    return any;
}
GCB_HD bool umi_is_duplex(const Umi &a, const Umi &b) {
    if (!umi_has_separator(a) || !umi_has_separator(b)) return false;
    UmiParts pa = umi_split(a), pb = umi_split(b);
    if (pa.n != 2 || pb.n != 2) return false;
    int a0 = pa.e0 - pa.b0, a1 = pa.e1 - pa.b1, c0 = pb.e0 - pb.b0, c1 = pb.e1 - pb.b1;
    if (a0 != c1 || a1 != c0) return false;
    for (int k = 0; k < a0; k++)
        if (umi_field(a, pa.b0 + k) != umi_field(b, pb.b1 + k)) return false;
    for (int k = 0; k < a1; k++)
        if (umi_field(a, pa.b1 + k) != umi_field(b, pb.b0 + k)) return false;
    return true;
}

// The two strand forms of a UMI for Cluster::isDuplex: `canon` = part 0, '_', part 1 and `swapped` = part 1, '_', part 0 when
// split() yields exactly two parts (returns false otherwise).  a and b are duplex partners iff canon(a) == swapped(b): the
// parts hold no '_', so the strings are equal iff the parts are, crosswise.  Computed once per family by duplex_kernel
// instead of two splits per candidate.
GCB_HD void umi_put(Umi &u, int k, int f) { u.w[k >> 4] |= (uint64_t)f << (60 - 4 * (k & 15)); }
GCB_HD bool umi_strand_forms(const Umi &u, Umi &canon, Umi &swapped) {
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) canon.w[k] = swapped.w[k] = 0ull;
    if (!umi_has_separator(u)) return false;
    const UmiParts p = umi_split(u);
    if (p.n != 2) return false;
    int k = 0;
    for (int q = p.b0; q < p.e0; q++) umi_put(canon, k++, umi_field(u, q));
    umi_put(canon, k++, UMI_UNDERSCORE);
    for (int q = p.b1; q < p.e1; q++) umi_put(canon, k++, umi_field(u, q));
    k = 0;
    for (int q = p.b1; q < p.e1; q++) umi_put(swapped, k++, umi_field(u, q));
    umi_put(swapped, k++, UMI_UNDERSCORE);
    for (int q = p.b0; q < p.e0; q++) umi_put(swapped, k++, umi_field(u, q));
    return true;
}

// ---- sequence helpers ------------------------------------------------------------------------
GCB_HD int base_at(const uint8_t *seq, int i) {  // bam_get_seq nibble order
    uint8_t b = seq[i >> 1];
    return (i & 1) ? (b & 0xF) : (b >> 4);
}
// BamUtil::fourbits2base (bamutil.cpp:149-165) as a class id: A C G T keep their code, the rest are 'N'
GCB_HD int base_letter(int code) { return (code == 1 || code == 2 || code == 4 || code == 8) ? code : 15; }

GCB_HD int sc8(int v) { return (int)(signed char)v; }  // the reference keeps scores in `char`

// Pair::qual2score, pair.cpp:77-86 (returns `char`)
GCB_HD int qual2score(const gcb_options &o, int q) {
    if (o.high_quality <= q) return sc8(o.score_high);
    if (o.moderate_quality <= q) return sc8(o.score_moderate);
    if (o.low_quality <= q) return sc8(o.score_low);
    return sc8(o.score_bad);
}

// record size of a read of l bases in a payload
GCB_HD int64_t record_bytes(int l) { return (int64_t)GCB_ALIGN4(l) + GCB_ALIGN4((l + 1) >> 1); }

#ifndef GCB_SIMT_CHECK
__device__ __forceinline__ void raise_error(int32_t *flag, int code) { atomicCAS(flag, 0, code); }
#else
inline void raise_error(int32_t *flag, int code) { if (*flag == 0) *flag = code; }
#endif

// ---- warp helpers ---------------------------------------------------------------------------------
GCB_DEV int lane_id() { return (int)(threadIdx.x & 31); }
GCB_DEV int warp_sum(int v) { return __reduce_add_sync(FULL, v); }
GCB_DEV int warp_min(int v) { return __reduce_min_sync(FULL, v); }
GCB_DEV int warp_max(int v) { return __reduce_max_sync(FULL, v); }
// ---- sub-warp groups: GS consecutive lanes (8, 16 or 32) that work on one cluster; every collective names the group's
// own lanes, so the groups of a warp are free to diverge
template <int GS>
struct Grp {
    static_assert(GS == 8 || GS == 16 || GS == 32, "group size");
    int gl;          // lane within the group
    int base;        // first lane of the group
    unsigned mask;   // the group's lanes
    GCB_DEV Grp() {
        const int l = lane_id();
        gl = l & (GS - 1);
        base = l & ~(GS - 1);
        mask = (GS == 32 ? 0xffffffffu : ((1u << GS) - 1u)) << base;
    }
    GCB_DEV int sum(int v) const { return __reduce_add_sync(mask, v); }
    GCB_DEV int min_of(int v) const { return __reduce_min_sync(mask, v); }
    GCB_DEV int max_of(int v) const { return __reduce_max_sync(mask, v); }
    GCB_DEV bool all(bool p) const { return __all_sync(mask, p) != 0; }
    GCB_DEV bool any(bool p) const { return __any_sync(mask, p) != 0; }
    GCB_DEV unsigned ballot(bool p) const { return __ballot_sync(mask, p) >> base; }  // bit k = lane k of the group
    template <typename T>
    GCB_DEV T shfl_xor(T v, int off) const { return __shfl_xor_sync(mask, v, off); }
    GCB_DEV void sync() const { __syncwarp(mask); }
};

GCB_DEV Umi umi_shfl(const Umi &u, int src) {
    Umi r;
#pragma unroll
    for (int k = 0; k < GCB_MAX_UMI_WORDS; k++) r.w[k] = __shfl_sync(FULL, u.w[k], src);
    return r;
}

}  // namespace gcb
