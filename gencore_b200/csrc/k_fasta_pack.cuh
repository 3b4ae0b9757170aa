// k_fasta_pack.cuh — FastaReader::readNext + to4bits (fastareader.cpp:58-152) for a whole FASTA text on the device: the
// reference genome is packed where it is going to be used.
//
// What the reference's reader does, restated per byte of the text after the first '>' (the constructor seeks it,
// fastareader.cpp:33-40):
//   * the reader alternates get(one char) / getline(rest).  A line whose get() char is '>' is a header: contig id = the
//     text up to the first space (fastareader.cpp:98-102), the rest of the line is dropped.
//   * every other get() char is appended RAW (only upper-cased): even a '\n' — a blank line — lands in the sequence
//     (as a non-ACGT base) and makes the getline swallow the NEXT physical line, whose first char is then filtered
//     like any other.  Line starts therefore alternate between "record starts" and swallowed lines
//     along a run of blank lines: a line start is a record start iff the run of '\n' bytes before it is odd.
//   * the rest of a line is filtered by str_keep_valid_sequence (util.h:194-210): letters, '-' and '*' stay.
//   * to4bits: A=1 T=2 C=3 G=4 other=0, EVEN index in the LOW nibble (fastareader.cpp:139-152).
// In parallel that is two scans over the text in blocks of FA_BLOCK bytes: a prefix maximum (the line start of every
// byte: is my line a header?) and a prefix sum of (emitted bases, headers) that gives every base its contig and its
// index in it.  Kernels: fa_anchor_kernel -> fa_scan_kernel<max> -> fa_count_kernel -> fa_scan_kernel<sum> ->
// fa_header_kernel -> (host: contig offsets) -> fa_pack_kernel.
#pragma once

#include "device_common.cuh"

namespace gcb {

constexpr int FA_THREADS = 256, FA_PER_THREAD = 16, FA_BLOCK = FA_THREADS * FA_PER_THREAD;

struct FaView {
    const uint8_t *t;  // the text from its first '>' on
    int64_t n;
};

GCB_DEV uint8_t fa_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
GCB_DEV bool fa_keep(uint8_t c) {  // str_keep_valid_sequence after upper-casing
    c = fa_upper(c);
    return (c >= 'A' && c <= 'Z') || c == '-' || c == '*';
}
GCB_DEV uint32_t fa_bits(uint8_t c) {  // FastaReader::base2bits
    c = fa_upper(c);
    return c == 'A' ? 1u : c == 'T' ? 2u : c == 'C' ? 3u : c == 'G' ? 4u : 0u;
}
GCB_DEV bool fa_line_start(const FaView &v, int64_t q) { return q == 0 || v.t[q - 1] == '\n'; }
// a line start where the reader does its get(): the run of '\n' before it is odd (position 0 is the first '>')
GCB_DEV bool fa_record_start(const FaView &v, int64_t q) {
    if (q == 0) return true;
    int64_t k = 0;
    while (q - 1 - k >= 0 && v.t[q - 1 - k] == '\n') k++;
    return (k & 1) != 0;
}
GCB_DEV bool fa_header_start(const FaView &v, int64_t q) { return v.t[q] == '>' && fa_line_start(v, q) && fa_record_start(v, q); }

// ---- block-wide inclusive scans of one value per thread (FA_THREADS threads)
struct FaMax { GCB_DEV static int64_t op(int64_t a, int64_t b) { return a > b ? a : b; } GCB_DEV static int64_t id() { return -1; } };
struct FaSum { GCB_DEV static int64_t op(int64_t a, int64_t b) { return a + b; } GCB_DEV static int64_t id() { return 0; } };
template <typename Op>
GCB_DEV int64_t fa_block_scan(int64_t v, int64_t *s_warp, int64_t &block_total) {
    const int lane = lane_id(), warp = (int)(threadIdx.x >> 5);
    for (int off = 1; off < WARP; off <<= 1) {
        const int64_t o = __shfl_up_sync(FULL, v, off);
        if (lane >= off) v = Op::op(v, o);
    }
    if (lane == WARP - 1) s_warp[warp] = v;
    __syncthreads();
    int64_t pre = Op::id(), tot = Op::id();
    for (int w = 0; w < FA_THREADS / WARP; w++) {
        if (w < warp) pre = Op::op(pre, s_warp[w]);
        tot = Op::op(tot, s_warp[w]);
    }
    __syncthreads();
    block_total = tot;
    return Op::op(pre, v);
}

// K1: the last line start of every block
__global__ void __launch_bounds__(FA_THREADS) fa_anchor_kernel(FaView v, int64_t *blk_anchor) {
    GCB_GRID_DEP();
    __shared__ int64_t s_warp[FA_THREADS / WARP];
    const int64_t q0 = (int64_t)blockIdx.x * FA_BLOCK + (int64_t)threadIdx.x * FA_PER_THREAD;
    int64_t a = -1;
    for (int k = 0; k < FA_PER_THREAD; k++)
        if (q0 + k < v.n && fa_line_start(v, q0 + k)) a = q0 + k;
    int64_t tot;
    fa_block_scan<FaMax>(a, s_warp, tot);
    if (threadIdx.x == 0) blk_anchor[blockIdx.x] = tot;
}

// the value of the thread before this one in an inclusive block scan (Op::id() for thread 0)
GCB_DEV int64_t fa_shift_right(int64_t incl, int64_t id) {
    __shared__ int64_t s_incl[FA_THREADS];
    s_incl[threadIdx.x] = incl;
    __syncthreads();
    const int64_t before = threadIdx.x == 0 ? id : s_incl[threadIdx.x - 1];
    __syncthreads();
    return before;
}

// K2 / K4: exclusive scan of the block aggregates, one CTA (nv values per block, interleaved); totals behind the last block
template <typename Op>
__global__ void __launch_bounds__(FA_THREADS) fa_scan_kernel(int64_t *blk, int64_t n_blocks, int nv) {
    GCB_GRID_DEP();
    __shared__ int64_t s_warp[FA_THREADS / WARP];
    for (int j = 0; j < nv; j++) {
        int64_t carry = Op::id();
        for (int64_t base = 0; base < n_blocks; base += FA_THREADS) {
            const int64_t i = base + threadIdx.x;
            const int64_t x = i < n_blocks ? blk[i * nv + j] : Op::id();
            int64_t tot;
            const int64_t incl = fa_block_scan<Op>(x, s_warp, tot);
            const int64_t excl = fa_shift_right(incl, Op::id());
            if (i < n_blocks) blk[i * nv + j] = Op::op(carry, excl);
            carry = Op::op(carry, tot);
        }
        if (threadIdx.x == 0) blk[n_blocks * nv + j] = carry;
    }
}

// what a thread knows about its FA_PER_THREAD bytes once the line starts are known
struct FaThread {
    uint32_t emit;    // bit k: byte k lands in the sequence
    uint32_t hdr;     // bit k: byte k is the '>' of a header line
    uint32_t bits[2]; // 2 x 8 nibbles: base2bits of the emitted bytes (0 elsewhere)
};
GCB_DEV FaThread fa_classify(const FaView &v, int64_t q0, int64_t anchor_in) {
    FaThread r;
    r.emit = r.hdr = 0u;
    r.bits[0] = r.bits[1] = 0u;
    int64_t anchor = anchor_in;  // line start of the byte before q0's (or of q0 itself below)
    bool in_hdr = false, have = false;
    for (int k = 0; k < FA_PER_THREAD; k++) {
        const int64_t q = q0 + k;
        if (q >= v.n) break;
        const uint8_t c = v.t[q];
        const bool ls = fa_line_start(v, q);
        bool rs = false;
        if (ls) {
            anchor = q;
            rs = fa_record_start(v, q);
            in_hdr = rs && c == '>';
            have = true;
            if (in_hdr) r.hdr |= 1u << k;
        } else if (!have) {  // the thread's first bytes continue a line that started before q0
            in_hdr = anchor >= 0 && v.t[anchor] == '>' && fa_record_start(v, anchor);
            have = true;
        }
        const bool e = !in_hdr && ((ls && rs) || fa_keep(c));
        if (e) {
            r.emit |= 1u << k;
            r.bits[k >> 3] |= fa_bits(c) << (4 * (k & 7));
        }
    }
    return r;
}
// the line start of the byte before q0 (-1: none), from the block's carry-in and the bytes of the block before q0
GCB_DEV int64_t fa_thread_anchor(const FaView &v, int64_t q0, int64_t blk_in, int64_t *s_warp) {
    int64_t a = -1;
    for (int k = 0; k < FA_PER_THREAD; k++)
        if (q0 + k < v.n && fa_line_start(v, q0 + k)) a = q0 + k;
    int64_t tot;
    const int64_t incl = fa_block_scan<FaMax>(a, s_warp, tot);
    return FaMax::op(blk_in, fa_shift_right(incl, -1));  // the maximum over the threads before this one
}

// K3: bases and headers of every block
__global__ void __launch_bounds__(FA_THREADS) fa_count_kernel(FaView v, const int64_t *blk_anchor_in, int64_t *blk_cnt) {
    GCB_GRID_DEP();
    __shared__ int64_t s_warp[FA_THREADS / WARP];
    const int64_t q0 = (int64_t)blockIdx.x * FA_BLOCK + (int64_t)threadIdx.x * FA_PER_THREAD;
    const int64_t anchor = fa_thread_anchor(v, q0, blk_anchor_in[blockIdx.x], s_warp);
    const FaThread f = fa_classify(v, q0, anchor);
    int64_t te, th;
    fa_block_scan<FaSum>(__popc(f.emit), s_warp, te);
    fa_block_scan<FaSum>(__popc(f.hdr), s_warp, th);
    if (threadIdx.x == 0) {
        blk_cnt[2 * (int64_t)blockIdx.x] = te;
        blk_cnt[2 * (int64_t)blockIdx.x + 1] = th;
    }
}

// K5: every header: where its line starts and how many bases precede it.  flag: 1 = more than max_contigs, 2 = a header the
// reference reads as something else (">\n", ">>", '>' as the last byte)
__global__ void __launch_bounds__(FA_THREADS) fa_header_kernel(FaView v, const int64_t *blk_anchor_in, const int64_t *blk_cnt_in, int64_t *hdr_pos,
                                                               int64_t *hdr_base, int32_t max_contigs, int32_t *flag) {
    GCB_GRID_DEP();
    __shared__ int64_t s_warp[FA_THREADS / WARP];
    const int64_t q0 = (int64_t)blockIdx.x * FA_BLOCK + (int64_t)threadIdx.x * FA_PER_THREAD;
    const int64_t anchor = fa_thread_anchor(v, q0, blk_anchor_in[blockIdx.x], s_warp);
    const FaThread f = fa_classify(v, q0, anchor);
    int64_t te, th;
    int64_t e = fa_block_scan<FaSum>(__popc(f.emit), s_warp, te) - __popc(f.emit) + blk_cnt_in[2 * (int64_t)blockIdx.x];
    int64_t h = fa_block_scan<FaSum>(__popc(f.hdr), s_warp, th) - __popc(f.hdr) + blk_cnt_in[2 * (int64_t)blockIdx.x + 1];
    for (int k = 0; k < FA_PER_THREAD; k++) {
        if ((f.hdr >> k) & 1u) {
            const int64_t q = q0 + k;
            if (h < max_contigs) {
                hdr_pos[h] = q;
                hdr_base[h] = e;
            } else {
                atomicOr(flag, 1);
            }
            if (q + 1 >= v.n || v.t[q + 1] == '\n' || v.t[q + 1] == '>') atomicOr(flag, 2);
            h++;
        }
        if ((f.emit >> k) & 1u) e++;
    }
}

// K6: the nibbles.  out32: the packed genome as 32-bit words, zeroed; contig_off: byte offset of every contig in it
__global__ void __launch_bounds__(FA_THREADS) fa_pack_kernel(FaView v, const int64_t *blk_anchor_in, const int64_t *blk_cnt_in, const int64_t *hdr_base,
                                                             const int64_t *contig_off, int32_t n_contigs, uint32_t *out32) {
    GCB_GRID_DEP();
    __shared__ int64_t s_warp[FA_THREADS / WARP];
    const int64_t q0 = (int64_t)blockIdx.x * FA_BLOCK + (int64_t)threadIdx.x * FA_PER_THREAD;
    const int64_t anchor = fa_thread_anchor(v, q0, blk_anchor_in[blockIdx.x], s_warp);
    const FaThread f = fa_classify(v, q0, anchor);
    int64_t te, th;
    int64_t e = fa_block_scan<FaSum>(__popc(f.emit), s_warp, te) - __popc(f.emit) + blk_cnt_in[2 * (int64_t)blockIdx.x];
    int64_t h = fa_block_scan<FaSum>(__popc(f.hdr), s_warp, th) - __popc(f.hdr) + blk_cnt_in[2 * (int64_t)blockIdx.x + 1];
    for (int k = 0; k < FA_PER_THREAD; k++) {
        if ((f.hdr >> k) & 1u) h++;
        if ((f.emit >> k) & 1u) {
            const int64_t cid = h - 1;  // (a base always follows its contig's header: the text starts with one)
            if (cid >= 0 && cid < n_contigs) {
                const int64_t idx = e - hdr_base[cid];
                const int64_t nib = 2 * contig_off[cid] + idx;  // nibble index in the packed genome: even index in the low nibble
                const uint32_t bits = (f.bits[k >> 3] >> (4 * (k & 7))) & 0xFu;
                if (bits) atomicOr(out32 + (nib >> 3), bits << (4 * (int)(nib & 7)));
            }
            e++;
        }
    }
}

}  // namespace gcb
