// gencore_b200_cli.cpp — the caller's side of the hot path: sorted BAM in, consensus BAM out,
// with the flags of the reference binary that reach the consensus path.  Host code only: it reads BAM records, keys them
// into clusters exactly as Gencore::addToProperCluster does (gencore.cpp:295-390, the 10 000-read tick flush, the
// watermark), packs the flushed clusters into gcb_batch, calls the engine through its C ABI (dlopen of
// libgencore_b200.so: UMI extraction, grouping, template selection, vote, duplex all run on the GPU) and replays the
// results the way Cluster::clusterByUMI's caller does (gencore.cpp:113-160: the ordered output set, write-behind the
// watermark).  There is no consensus arithmetic in this file and no CPU fallback: without the engine library and a
// B200 it exits.  Not built here: Stats, JSON/HTML reports, BED depth (out of scope), SAM text I/O.
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <set>
#include <string>
#include <vector>

#include "gencore_b200.h"

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
const bool g_timing = getenv("GCB_TIMING") != nullptr;  // per-phase wall times on stderr (and when the phase ended, since the start of main)
const double g_t_start = now_s();
void lap(const char *what, double &t0) {
    if (!g_timing) return;
    const double t = now_s();
    fprintf(stderr, "[timing] %-28s %8.3f s   (at %7.3f)\n", what, t - t0, t - g_t_start);
    t0 = t;
}

[[noreturn]] void die(const std::string &msg) {  // util.h:250-253 error_exit
    fprintf(stderr, "ERROR: %s\n", msg.c_str());
    exit(-1);
}

// ---------------------------------------------------------------------------------------------- BGZF
// Both directions are block-parallel (the codec, not the GPU, bounds the whole tool): a pool of
// worker threads inflates / deflates 64 KiB blocks, the blocks are consumed / written in file order.
const size_t BGZF_PAYLOAD = 0xff00;
const int CODEC_SLOTS = 128;

struct CodecTask {
    std::vector<uint8_t> in, out;
    uint32_t isize = 0;
    bool ready = false;  // `out` is complete
};

struct WorkerPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<std::pair<CodecTask *, bool>> work;  // (task, inflate?)
    bool stop = false;
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; i++) threads.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : threads) t.join();
    }
    static void inflate_block(CodecTask *t) {
        t->out.resize(t->isize);
        if (!t->isize) return;
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) die("zlib");
        zs.next_in = t->in.data();
        zs.avail_in = (uInt)t->in.size();
        zs.next_out = t->out.data();
        zs.avail_out = t->isize;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) die("corrupt BGZF block");
    }
    static void deflate_block(CodecTask *t) {
        const size_t n = t->in.size();
        t->out.resize(18 + n + 1024 + 8);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) die("zlib");
        zs.next_in = t->in.data();
        zs.avail_in = (uInt)n;
        zs.next_out = t->out.data() + 18;
        zs.avail_out = (uInt)(t->out.size() - 26);
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) die("deflate failed");
        const size_t total = 18 + clen + 8;
        const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)((total - 1) & 0xff), (uint8_t)((total - 1) >> 8)};
        memcpy(t->out.data(), hdr, 18);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), t->in.data(), (uInt)n);
        const uint8_t tail[8] = {(uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24),
                                 (uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        memcpy(t->out.data() + 18 + clen, tail, 8);
        t->out.resize(total);
    }
    void loop() {
        for (;;) {
            std::pair<CodecTask *, bool> job;
            {
                std::unique_lock<std::mutex> l(mu);
                cv_work.wait(l, [this] { return stop || !work.empty(); });
                if (work.empty()) return;
                job = work.front();
                work.pop_front();
            }
            if (job.second) inflate_block(job.first);
            else deflate_block(job.first);
            {
                std::lock_guard<std::mutex> l(mu);
                job.first->ready = true;
            }
            cv_done.notify_all();
        }
    }
    void submit(CodecTask *t, bool inflate) {
        {
            std::lock_guard<std::mutex> l(mu);
            t->ready = false;
            work.emplace_back(t, inflate);
        }
        cv_work.notify_one();
    }
    void wait(CodecTask *t) {
        std::unique_lock<std::mutex> l(mu);
        cv_done.wait(l, [t] { return t->ready; });
    }
};

struct BgzfReader {
    FILE *fp = nullptr;
    WorkerPool *pool = nullptr;
    std::vector<CodecTask> slots = std::vector<CodecTask>(CODEC_SLOTS);
    size_t head = 0, tail = 0;  // blocks [head, tail) are in flight; slot = index % CODEC_SLOTS
    bool file_eof = false;
    std::vector<uint8_t> buf;
    size_t pos = 0;
    bool read_raw_block(CodecTask &t) {  // the compressed bytes of the next block
        uint8_t hdr[12];
        const size_t got = fread(hdr, 1, 12, fp);
        if (got == 0) return false;
        if (got != 12 || hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) die("input is not BGZF");
        const unsigned xlen = hdr[10] | (hdr[11] << 8);
        std::vector<uint8_t> extra(xlen);
        if (fread(extra.data(), 1, xlen, fp) != xlen) die("truncated BGZF block");
        int bsize = -1;
        for (size_t i = 0; i + 4 <= extra.size();) {
            const unsigned slen = extra[i + 2] | (extra[i + 3] << 8);
            if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2) bsize = extra[i + 4] | (extra[i + 5] << 8);
            i += 4 + slen;
        }
        const long clen = (long)bsize + 1 - 12 - (long)xlen - 8;
        if (bsize < 0 || clen < 0) die("bad BGZF block header");
        t.in.resize((size_t)clen);
        uint8_t tail8[8];
        if ((clen && fread(t.in.data(), 1, (size_t)clen, fp) != (size_t)clen) || fread(tail8, 1, 8, fp) != 8) die("truncated BGZF block");
        t.isize = tail8[4] | (tail8[5] << 8) | (tail8[6] << 16) | ((uint32_t)tail8[7] << 24);
        return true;
    }
    void fill() {  // keep the pool busy
        while (!file_eof && tail - head < (size_t)CODEC_SLOTS) {
            CodecTask &t = slots[tail % CODEC_SLOTS];
            if (!read_raw_block(t)) {
                file_eof = true;
                break;
            }
            pool->submit(&t, true);
            tail++;
        }
    }
    bool next_block() {
        fill();
        if (head == tail) return false;
        CodecTask &t = slots[head % CODEC_SLOTS];
        pool->wait(&t);
        buf.swap(t.out);
        pos = 0;
        head++;
        return true;
    }
    size_t read(void *dst, size_t n) {
        uint8_t *out = (uint8_t *)dst;
        size_t done = 0;
        while (done < n) {
            if (pos == buf.size()) {
                if (!next_block()) break;
                continue;
            }
            const size_t take = std::min(buf.size() - pos, n - done);
            memcpy(out + done, buf.data() + pos, take);
            pos += take;
            done += take;
        }
        return done;
    }
};

struct BgzfWriter {
    FILE *fp = nullptr;
    WorkerPool *pool = nullptr;
    std::vector<CodecTask> slots = std::vector<CodecTask>(CODEC_SLOTS);
    size_t head = 0, tail = 0;
    std::vector<uint8_t> pending;
    void drain(size_t keep) {  // write finished blocks in order until at most `keep` are in flight
        while (tail - head > keep) {
            CodecTask &t = slots[head % CODEC_SLOTS];
            pool->wait(&t);
            if (fwrite(t.out.data(), 1, t.out.size(), fp) != t.out.size()) die("Writing failed, exiting ...");
            head++;
        }
    }
    void emit(const uint8_t *src, size_t n) {
        drain(CODEC_SLOTS - 1);
        CodecTask &t = slots[tail % CODEC_SLOTS];
        t.in.assign(src, src + n);
        pool->submit(&t, false);
        tail++;
    }
    void flush(bool all) {
        size_t off = 0;
        while (pending.size() - off >= BGZF_PAYLOAD || (all && off < pending.size())) {
            const size_t n = std::min(pending.size() - off, BGZF_PAYLOAD);
            emit(pending.data() + off, n);
            off += n;
        }
        pending.erase(pending.begin(), pending.begin() + (long)off);
    }
    void write(const void *src, size_t n) {
        const uint8_t *p = (const uint8_t *)src;
        pending.insert(pending.end(), p, p + n);
        if (pending.size() >= 16 * BGZF_PAYLOAD) flush(false);
    }
    void close() {
        flush(true);
        drain(0);
        static const uint8_t eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        fwrite(eof_block, 1, 28, fp);
        fclose(fp);
    }
};

// ---------------------------------------------------------------------------------------------- BAM records
struct Rec {
    int32_t tid = -1, pos = -1, mtid = -1, mpos = -1, isize = 0, l_seq = 0;
    uint16_t bin = 0, flag = 0;
    uint8_t mapq = 0;
    std::string qname;            // without the NUL
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> seq, qual, aux;
    uint64_t serial = 0;          // stands in for the heap address that breaks ties in gencore.h:19-47
    int l_qname_padded() const { return (int)((qname.size() + 1 + 3) & ~(size_t)3); }  // htslib's in-memory l_qname (Q14)
};

uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
void put32(std::vector<uint8_t> &v, uint32_t x) {
    for (int k = 0; k < 4; k++) v.push_back((uint8_t)(x >> (8 * k)));
}

struct BamHeader {
    std::string text;
    std::vector<std::string> names;
    std::vector<int32_t> lens;
};

BamHeader read_header(BgzfReader &in) {
    uint8_t b4[4];
    BamHeader h;
    if (in.read(b4, 4) != 4 || memcmp(b4, "BAM\1", 4) != 0) die("input is not a BAM file");
    if (in.read(b4, 4) != 4) die("truncated BAM header");
    h.text.resize(le32(b4));
    if (in.read(&h.text[0], h.text.size()) != h.text.size()) die("truncated BAM header");
    if (in.read(b4, 4) != 4) die("truncated BAM header");
    const uint32_t n = le32(b4);
    for (uint32_t i = 0; i < n; i++) {
        if (in.read(b4, 4) != 4) die("truncated BAM header");
        std::string name(le32(b4), '\0');
        if (in.read(&name[0], name.size()) != name.size() || in.read(b4, 4) != 4) die("truncated BAM header");
        if (!name.empty() && name.back() == '\0') name.pop_back();
        h.names.push_back(name);
        h.lens.push_back((int32_t)le32(b4));
    }
    return h;
}

void write_header(BgzfWriter &out, const BamHeader &h) {
    std::vector<uint8_t> v = {'B', 'A', 'M', 1};
    put32(v, (uint32_t)h.text.size());
    v.insert(v.end(), h.text.begin(), h.text.end());
    put32(v, (uint32_t)h.names.size());
    for (size_t i = 0; i < h.names.size(); i++) {
        put32(v, (uint32_t)h.names[i].size() + 1);
        v.insert(v.end(), h.names[i].begin(), h.names[i].end());
        v.push_back(0);
        put32(v, (uint32_t)h.lens[i]);
    }
    out.write(v.data(), v.size());
}

// returns false at EOF
bool read_record(BgzfReader &in, Rec &r) {
    uint8_t b4[4];
    const size_t got = in.read(b4, 4);
    if (got == 0) return false;
    if (got != 4) die("truncated BAM record");
    const uint32_t block = le32(b4);
    static thread_local std::vector<uint8_t> d;  // (one buffer for every record a thread reads: no allocation, no zero fill)
    if (d.size() < block) d.resize(block);
    if (block < 32 || in.read(d.data(), block) != block) die("truncated BAM record");
    r.tid = (int32_t)le32(&d[0]);
    r.pos = (int32_t)le32(&d[4]);
    const unsigned l_name = d[8];
    r.mapq = d[9];
    r.bin = (uint16_t)(d[10] | (d[11] << 8));
    const unsigned n_cig = d[12] | (d[13] << 8);
    r.flag = (uint16_t)(d[14] | (d[15] << 8));
    r.l_seq = (int32_t)le32(&d[16]);
    r.mtid = (int32_t)le32(&d[20]);
    r.mpos = (int32_t)le32(&d[24]);
    r.isize = (int32_t)le32(&d[28]);
    size_t off = 32;
    if (r.l_seq < 0) die("corrupt BAM record");  // (a negative length would wrap the size check below)
    if (off + l_name + 4 * n_cig + (size_t)(r.l_seq + 1) / 2 + (size_t)r.l_seq > block) die("corrupt BAM record");
    r.qname.assign((const char *)&d[off], l_name ? strnlen((const char *)&d[off], l_name) : 0);
    off += l_name;
    r.cigar.resize(n_cig);
    for (unsigned k = 0; k < n_cig; k++) r.cigar[k] = le32(&d[off + 4 * k]);
    off += 4 * n_cig;
    r.seq.assign(d.begin() + (long)off, d.begin() + (long)(off + (size_t)(r.l_seq + 1) / 2));
    off += (size_t)(r.l_seq + 1) / 2;
    r.qual.assign(d.begin() + (long)off, d.begin() + (long)(off + (size_t)r.l_seq));
    off += (size_t)r.l_seq;
    r.aux.assign(d.begin() + (long)off, d.begin() + (long)block);
    return true;
}

void write_record(BgzfWriter &out, const Rec &r) {
    std::vector<uint8_t> v;
    const uint32_t l_name = (uint32_t)r.qname.size() + 1;
    const uint32_t block = 32 + l_name + 4 * (uint32_t)r.cigar.size() + (uint32_t)r.seq.size() + (uint32_t)r.qual.size() + (uint32_t)r.aux.size();
    v.reserve(block + 4);
    put32(v, block);
    put32(v, (uint32_t)r.tid);
    put32(v, (uint32_t)r.pos);
    v.push_back((uint8_t)l_name);
    v.push_back(r.mapq);
    v.push_back((uint8_t)r.bin);
    v.push_back((uint8_t)(r.bin >> 8));
    v.push_back((uint8_t)r.cigar.size());
    v.push_back((uint8_t)(r.cigar.size() >> 8));
    v.push_back((uint8_t)r.flag);
    v.push_back((uint8_t)(r.flag >> 8));
    put32(v, (uint32_t)r.l_seq);
    put32(v, (uint32_t)r.mtid);
    put32(v, (uint32_t)r.mpos);
    put32(v, (uint32_t)r.isize);
    v.insert(v.end(), r.qname.begin(), r.qname.end());
    v.push_back(0);
    for (uint32_t c : r.cigar) put32(v, c);
    v.insert(v.end(), r.seq.begin(), r.seq.end());
    v.insert(v.end(), r.qual.begin(), r.qual.end());
    v.insert(v.end(), r.aux.begin(), r.aux.end());
    out.write(v.data(), v.size());
}

// bam_aux_get: pointer to the type byte of `tag`, or null
const uint8_t *aux_find(const std::vector<uint8_t> &aux, const char tag[2], size_t *index = nullptr) {
    size_t i = 0;
    while (i + 3 <= aux.size()) {
        const bool hit = aux[i] == (uint8_t)tag[0] && aux[i + 1] == (uint8_t)tag[1];
        const char type = (char)aux[i + 2];
        if (hit) {
            if (index) *index = i + 2;
            return &aux[i + 2];
        }
        i += 3;
        switch (type) {
            case 'A': case 'c': case 'C': i += 1; break;
            case 's': case 'S': i += 2; break;
            case 'i': case 'I': case 'f': i += 4; break;
            case 'Z': case 'H':
                while (i < aux.size() && aux[i]) i++;
                i++;
                break;
            case 'B': {
                if (i + 5 > aux.size()) return nullptr;
                const char sub = (char)aux[i];
                const uint32_t n = le32(&aux[i + 1]);
                const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                i += 5 + w * n;
                break;
            }
            default: return nullptr;
        }
    }
    return nullptr;
}

// ---------------------------------------------------------------------------------------------- options (main.cpp:29-89)
struct Cli {
    std::string input = "-", output = "-", ref, umi_prefix = "auto", engine;
    gcb_options opt;
    int umi_diff_threshold = 1;  // -d properReadsUmiDiffThreshold; unproper reads use 0 (options.cpp:12-13)
    int device = 0;
    // (not reference flags) one input over several processes / GPUs: `--shard i/N` keeps only the clusters whose (contig, left)
    // falls into the i-th of N equal windows of the concatenated contigs; `--merge` joins the shards' outputs
    int shard_index = 0, shard_count = 1;
    std::vector<std::string> merge_inputs;
};

// ---------------------------------------------------------------------------------------------- the engine (C ABI, dlopen)
struct Engine {
    void *lib = nullptr;
    gcb_ctx *ctx = nullptr;
    decltype(&gcb_create) create;
    decltype(&gcb_destroy) destroy;
    decltype(&gcb_last_error) last_error;
    decltype(&gcb_set_reference) set_reference;
    decltype(&gcb_consensus_batch) consensus_batch;
    decltype(&gcb_extract_umi) extract_umi;
    decltype(&gcb_pack_fasta) pack_fasta;
    decltype(&gcb_default_options) default_options;
    template <typename F>
    void sym(F &f, const char *name) {
        f = (F)dlsym(lib, name);
        if (!f) die(std::string("engine library lacks ") + name);
    }
    void open(const std::string &path) {
        lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!lib) die(std::string("cannot load the engine library (there is no CPU fallback): ") + dlerror());
        sym(create, "gcb_create");
        sym(destroy, "gcb_destroy");
        sym(last_error, "gcb_last_error");
        sym(set_reference, "gcb_set_reference");
        sym(consensus_batch, "gcb_consensus_batch");
        sym(extract_umi, "gcb_extract_umi");
        sym(pack_fasta, "gcb_pack_fasta");
        sym(default_options, "gcb_default_options");
    }
    void check(int rc, const char *what) {
        if (rc != GCB_OK) die(std::string(what) + ": " + (ctx ? last_error(ctx) : "engine error") + " (status " + std::to_string(rc) + ")");
    }
};

// ---------------------------------------------------------------------------------------------- FASTA (fastareader.cpp:58-152)
struct Genome {
    std::vector<uint8_t> packed;
    std::vector<int64_t> off, len;
    std::map<std::string, int> index;
};

// The file is read here; reading it the way FastaReader does (readNext + to4bits) and packing it is the engine's work
// (gcb_pack_fasta: the genome is packed on the GPU).  Of two contigs with one id the later wins, as in the reference's map.
Genome load_fasta(const std::string &path, Engine &eng) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) die("Failed to open file: " + path);
    std::string all;
    char chunk[1 << 16];
    size_t n;
    while ((n = fread(chunk, 1, sizeof chunk, fp)) > 0) all.append(chunk, n);
    fclose(fp);
    Genome g;
    int32_t max_contigs = 1;
    for (char c : all) max_contigs += c == '>';
    g.packed.assign(all.size() / 2 + 16 * (size_t)max_contigs + 16, 0);
    g.off.assign((size_t)max_contigs, 0);
    g.len.assign((size_t)max_contigs, 0);
    std::vector<int64_t> name_off((size_t)max_contigs);
    std::vector<int32_t> name_len((size_t)max_contigs);
    int32_t nc = 0;
    int64_t bytes = 0;
    eng.check(eng.pack_fasta(eng.ctx, all.data(), (int64_t)all.size(), max_contigs, g.packed.data(), (int64_t)g.packed.size(), g.off.data(), g.len.data(),
                             name_off.data(), name_len.data(), &nc, &bytes),
              "gcb_pack_fasta");
    g.packed.resize((size_t)bytes);
    g.off.resize((size_t)nc);
    g.len.resize((size_t)nc);
    for (int32_t i = 0; i < nc; i++) g.index[all.substr((size_t)name_off[(size_t)i], (size_t)name_len[(size_t)i])] = i;
    return g;
}

// ---------------------------------------------------------------------------------------------- clusters and the event log
struct PairRec {  // Pair (pair.h:12-68): the records only
    Rec *left = nullptr, *right = nullptr;
};
typedef std::map<std::string, PairRec> PairMap;  // Cluster::mPairs, keyed by qname + htslib's NUL padding (bamutil.cpp:19-21)

struct ClusterJob {
    int tid;
    long right;
    int thr;
    bool passthrough;  // finishConsensus writes clusters with a negative coordinate as they are (gencore.cpp:400-407)
    std::vector<PairRec> pairs;  // map order
};

struct Event {
    enum Kind { PASS, CLUSTERS, CLEAR_OUTSET } kind;
    Rec *rec = nullptr;                // PASS: outputBam(b, true) of a read whose mate is unmapped (gencore.cpp:307-309)
    std::vector<ClusterJob> jobs;      // CLUSTERS: one tick flush or finishConsensus
    bool set_watermark = false;        // the flush updated mProcessedTid / mProcessedPos (gencore.cpp:386-389)
    int wm_tid = -1, wm_pos = -1;
};

struct OutComp {  // gencore.h:19-47, the serial number in place of the data pointer
    bool operator()(const Rec *a, const Rec *b) const {
        if (a->tid >= 0) {
            if (b->tid < 0) return true;
            if (b->tid != a->tid) return b->tid > a->tid;
            if (b->pos != a->pos) return b->pos > a->pos;
            if (b->mtid != a->mtid) return b->mtid > a->mtid;
            if (b->mpos != a->mpos) return b->mpos > a->mpos;
            if (b->isize != a->isize) return b->isize > a->isize;
            return b->serial > a->serial;
        }
        if (b->tid < 0) return b->serial > a->serial;
        return false;
    }
};

struct Pipeline {
    Cli cli;
    Engine eng;
    BamHeader hdr;
    BgzfWriter out;
    Genome genome;
    std::vector<int> tid_to_contig;
    std::map<int, std::map<int, std::map<long, PairMap>>> clusters;  // mProperClusters (gencore.h)
    std::vector<Event> log;
    size_t pending_pairs = 0;
    std::set<Rec *, OutComp> out_set;
    int processed_tid = -1, processed_pos = -1;  // mProcessedTid / mProcessedPos
    bool out_set_cleared = false, clusters_finished = false;
    uint64_t serial = 0;
    long tick = 0;
    std::string prefix;
    std::future<void> engine_ready;
    std::vector<int64_t> lin_off;  // --shard: first coordinate of every contig in the concatenation of the header's contigs
    int64_t lin_total = 0;

    // --shard i/N: a cluster belongs to the window its (contig, left) falls into, a read whose mate is unmapped to the window of
    // its own position.  Clusters are independent (gencore.cpp:76, cluster.cpp:55), and whether one is flushed at a tick depends
    // only on its own key and on the read that ticks (see add_to_proper_cluster), so a process that COUNTS every clustered read
    // of the input but KEEPS only its own window's reads flushes its clusters at the same ticks, with the same UMI threshold
    // (Q18), as the unsharded run.
    bool mine(int tid, int left) const {
        if (cli.shard_count <= 1) return true;
        const int64_t len = hdr.lens[(size_t)tid];
        const int64_t x = lin_off[(size_t)tid] + std::min<int64_t>(std::max<int64_t>(left, 0), std::max<int64_t>(len - 1, 0));
        const int owner = lin_total > 0 ? (int)std::min<int64_t>(x * cli.shard_count / lin_total, cli.shard_count - 1) : 0;
        return owner == cli.shard_index;
    }

    // ---- output side (gencore.cpp:83-160)
    void write_bam(Rec *r) { write_record(out, *r); }
    void output_bam(Rec *r, bool is_left) {
        auto ret = out_set.insert(r);
        auto insertpos = ret.first;
        ++insertpos;
        if (is_left) {
            auto it = out_set.begin();
            for (; it != insertpos; ++it) {
                if (processed_pos == -1 || (*it)->tid > processed_tid || ((*it)->tid == processed_tid && (*it)->pos >= processed_pos)) break;
                write_bam(*it);
                delete *it;
            }
            out_set.erase(out_set.begin(), it);
        }
    }
    void output_pair(Rec *left, Rec *right) {
        if (left) output_bam(left, true);
        if (right) output_bam(right, false);
    }
    void output_out_set() {
        for (Rec *r : out_set) {
            write_bam(r);
            delete r;
        }
        out_set.clear();  // (mOutSetCleared is the reading side's flag: Pipeline::run sets it when it logs the event)
    }

    // ---- input side
    static std::string map_key(const Rec &r) {  // BamUtil::getQName: the name with htslib's padding NULs
        std::string k = r.qname;
        k.resize((size_t)r.l_qname_padded(), '\0');
        return k;
    }
    void add_read(PairMap &pm, Rec *r) {  // Cluster::addRead, cluster.cpp:260-273
        PairRec fresh;
        fresh.left = r;
        auto ins = pm.emplace(map_key(*r), fresh);  // (one key, one walk of the map)
        if (!ins.second) {
            delete ins.first->second.right;  // Pair::setRight destroys a previous mRight (Q25)
            ins.first->second.right = r;
        }
    }
    static void take_jobs(std::vector<ClusterJob> &jobs, int tid, long right, int thr, bool passthrough, PairMap &pm) {
        ClusterJob j;
        j.tid = tid;
        j.right = right;
        j.thr = thr;
        j.passthrough = passthrough;
        for (auto &kv : pm) j.pairs.push_back(kv.second);
        jobs.push_back(std::move(j));
    }
    void add_to_proper_cluster(Rec *b) {  // gencore.cpp:295-390
        const int tid = b->tid, bpos = b->pos;
        int left = b->pos;
        long right;
        if (b->mtid == b->tid && abs(b->mpos - b->pos) < 100000) {
            if (b->isize < 0) left = b->mpos;
            right = (long)left + abs(b->isize) - 1;
        } else {
            if (b->mtid < 0) {
                if (!mine(tid, b->pos)) {
                    delete b;
                    return;
                }
                Event e;
                e.kind = Event::PASS;
                e.rec = b;
                log.push_back(std::move(e));
                return;
            }
            right = -1L * (long)hdr.lens[(size_t)b->tid] * (long)(b->mtid + 1) + (long)b->mpos;
        }
        if (mine(tid, left)) add_read(clusters[tid][left][right], b);
        else delete b;  // (another shard's read: it still counts for the tick)
        tick++;
        if (tick % 10000 != 0) return;
        // the tick flush: every cluster the sorted input can no longer add to
        Event e;
        e.kind = Event::CLUSTERS;
        bool need_break = false;
        int cur_tid = 0x7FFFFFFF, cur_pos = -1, processed = 0;
        for (auto i1 = clusters.begin(); i1 != clusters.end();) {
            if (i1->first > tid || need_break) {
                if (cur_tid > i1->first) {
                    cur_tid = i1->first;
                    cur_pos = processed;
                }
                break;
            }
            processed = hdr.lens[(size_t)i1->first];
            for (auto i2 = i1->second.begin(); i2 != i1->second.end();) {
                if (i1->first == tid && i2->first >= bpos) {
                    if (processed > i2->first) processed = i2->first;
                    need_break = true;
                    break;
                }
                for (auto i3 = i2->second.begin(); i3 != i2->second.end();) {
                    if (i1->first == tid && i3->first >= bpos) break;
                    take_jobs(e.jobs, i1->first, i3->first, cli.umi_diff_threshold, false, i3->second);
                    i3 = i2->second.erase(i3);
                }
                if (i2->second.empty()) {
                    i2 = i1->second.erase(i2);
                } else {
                    if (processed > i2->first) processed = i2->first;
                    ++i2;
                }
            }
            if (i1->second.empty()) {
                i1 = clusters.erase(i1);
                cur_pos = processed;
            } else {
                if (cur_tid > i1->first) {
                    cur_tid = i1->first;
                    cur_pos = processed;
                }
                ++i1;
            }
        }
        if (cur_tid != 0x7FFFFFFF) {
            e.set_watermark = true;
            e.wm_tid = cur_tid;
            e.wm_pos = cur_pos;
        }
        for (const ClusterJob &j : e.jobs) pending_pairs += j.pairs.size();
        log.push_back(std::move(e));
    }
    void finish_consensus() {  // gencore.cpp:392-434: everything left, UMI threshold 0
        Event e;
        e.kind = Event::CLUSTERS;
        for (auto &k1 : clusters)
            for (auto &k2 : k1.second)
                for (auto &k3 : k2.second) take_jobs(e.jobs, k1.first, k3.first, 0, k1.first < 0 || k2.first < 0, k3.second);
        clusters.clear();
        for (const ClusterJob &j : e.jobs) pending_pairs += j.pairs.size();
        log.push_back(std::move(e));
    }

    // ---- the engine call over every CLUSTERS event of a batch of the log, then the replay of that batch.  Three threads: the
    // one that reads and keys (clusters, tick, log, pending_pairs, serial, out_set_cleared) hands whole batches of the log to
    // the one that packs them and calls the engine (eng, genome, prefix), which hands the batch and its results to the one that
    // replays and writes (out_set, watermark, out).  Batches go through in order, so the sequence of outputs is that of one
    // thread; the next 200 000 pairs are parsed and clustered while the engine works on the current ones and the writer on
    // the ones before.
    struct Packed {  // a batch of the log after the engine call: what the replay needs
        std::vector<Event> log;
        std::vector<int32_t> cpo, n_groups;
        std::vector<gcb_group_result> groups;
        std::vector<uint8_t> out_payload;
        std::vector<uint64_t> umi;
        std::vector<const Rec *> slot_rec;  // read slot -> record
        int umi_words = 2;
        int32_t n_pairs = 0;
        // the packed batch (gencore_b200.h "Encoding conventions"), used by the engine stage only
        std::vector<int32_t> ctid;  // the clusters' tids (the engine stage maps them to contigs of the FASTA once that is loaded)
        std::vector<uint8_t> cflags;
        std::vector<gcb_read_desc> reads;
        std::vector<uint32_t> cigar;
        std::vector<uint8_t> payload;
        std::string names;
        std::vector<int64_t> name_off;
    };
    template <typename T>
    struct Channel {  // a bounded queue between two stages (a batch holds its reads)
        std::mutex mu;
        std::condition_variable cv;
        std::deque<T> q;
        size_t cap = 2;
        bool closed = false;
        void push(T &&v) {
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [this] { return q.size() < cap; });
                q.push_back(std::move(v));
            }
            cv.notify_all();
        }
        bool pop(T &v) {
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [this] { return !q.empty() || closed; });
                if (q.empty()) return false;
                v = std::move(q.front());
                q.pop_front();
            }
            cv.notify_all();
            return true;
        }
        void close() {
            {
                std::lock_guard<std::mutex> l(mu);
                closed = true;
            }
            cv.notify_all();
        }
    };
    Channel<std::vector<Event>> to_pack;
    Channel<std::unique_ptr<Packed>> to_engine, to_writer;
    void submit_log() {
        std::vector<Event> batch;
        batch.swap(log);
        pending_pairs = 0;
        to_pack.push(std::move(batch));
    }
    void pack_stage() {
        std::vector<Event> batch;
        while (to_pack.pop(batch)) to_engine.push(pack(std::move(batch)));
        to_engine.close();
    }
    void engine_stage() {
        std::unique_ptr<Packed> pk;
        while (to_engine.pop(pk)) {
            run_engine(*pk);
            to_writer.push(std::move(pk));
        }
        to_writer.close();
    }
    void writer_stage() {
        std::unique_ptr<Packed> pk;
        while (to_writer.pop(pk)) replay(*pk);
    }
    std::unique_ptr<Packed> pack(std::vector<Event> &&batch);
    void run_engine(Packed &pk);
    void replay(Packed &pk);
    void run();
};

// 1. pack (gencore_b200.h "Encoding conventions"): needs no engine, so it runs while the CUDA context starts up
std::unique_ptr<Pipeline::Packed> Pipeline::pack(std::vector<Event> &&batch) {
    double t0 = now_s();
    std::unique_ptr<Packed> pkp(new Packed());
    Packed &pk = *pkp;
    pk.log = std::move(batch);
    std::vector<Event> &log = pk.log;
    std::vector<int32_t> &cpo = pk.cpo, &ctid = pk.ctid;
    cpo.push_back(0);
    std::vector<uint8_t> &cflags = pk.cflags;
    std::vector<gcb_read_desc> &reads = pk.reads;
    std::vector<uint32_t> &cigar = pk.cigar;
    std::vector<uint8_t> &payload = pk.payload;
    std::vector<const Rec *> &slot_rec = pk.slot_rec;
    std::string &names = pk.names;
    std::vector<int64_t> &name_off = pk.name_off;
    name_off.push_back(0);
    for (Event &e : log) {
        if (e.kind != Event::CLUSTERS) continue;
        for (ClusterJob &j : e.jobs) {
            if (j.passthrough) continue;
            payload.resize((payload.size() + 15) & ~(size_t)15, 0);
            for (const PairRec &p : j.pairs) {
                for (int side = 0; side < 2; side++) {
                    const Rec *r = side == 0 ? p.left : p.right;
                    gcb_read_desc d;
                    memset(&d, 0, sizeof d);
                    d.l_qseq = -1;
                    if (r) {
                        d.data_off = (int64_t)payload.size();
                        d.l_qseq = r->l_seq;
                        d.pos = r->pos;
                        d.isize = r->isize;
                        d.cigar_off = (int32_t)cigar.size();
                        d.n_cigar = (uint16_t)r->cigar.size();
                        d.l_qname = (uint16_t)r->l_qname_padded();
                        cigar.insert(cigar.end(), r->cigar.begin(), r->cigar.end());
                        payload.insert(payload.end(), r->qual.begin(), r->qual.end());
                        payload.resize((payload.size() + 3) & ~(size_t)3, 0);
                        payload.insert(payload.end(), r->seq.begin(), r->seq.end());
                        payload.resize((payload.size() + 3) & ~(size_t)3, 0);
                        // the UMI comes from the MI:Z tag when there is one, else from the name (bamutil.cpp:23-38)
                        const uint8_t *mi = aux_find(r->aux, "MI");
                        if (mi && (*mi == 'Z' || *mi == 'H')) {  // (the value ends at its NUL or at the end of the aux block)
                            const char *v = (const char *)mi + 1;
                            names.append(v, strnlen(v, (size_t)(r->aux.data() + r->aux.size() - (const uint8_t *)v)));
                        } else {
                            names.append(r->qname);
                        }
                    }
                    name_off.push_back((int64_t)names.size());
                    reads.push_back(d);
                    slot_rec.push_back(r);
                }
            }
            cpo.push_back((int32_t)(reads.size() / 2));
            ctid.push_back(j.tid);
            cflags.push_back((uint8_t)((j.right < 0 ? GCB_CLUSTER_CROSS_CONTIG : 0) | (j.thr << GCB_CLUSTER_UMI_THR_SHIFT)));
        }
    }
    payload.resize((payload.size() + 15) & ~(size_t)15, 0);
    pk.n_pairs = (int32_t)(reads.size() / 2);
    lap("pack batch", t0);
    return pkp;
}

// 2.-3. the engine over a packed batch: UMIs, then Cluster::clusterByUMI for every cluster
void Pipeline::run_engine(Packed &pk) {
    double t0 = now_s();
    if (engine_ready.valid()) {
        engine_ready.get();
        lap("wait for engine start-up", t0);
        for (const std::string &n : hdr.names) {
            auto it = genome.index.find(n);
            tid_to_contig.push_back(it == genome.index.end() ? -1 : it->second);
        }
    }
    std::vector<int32_t> &cpo = pk.cpo, cref;
    for (int32_t t : pk.ctid) cref.push_back(t >= 0 && (size_t)t < tid_to_contig.size() ? tid_to_contig[(size_t)t] : -1);
    std::vector<uint8_t> &cflags = pk.cflags;
    std::vector<gcb_read_desc> &reads = pk.reads;
    std::vector<uint32_t> &cigar = pk.cigar;
    std::vector<uint8_t> &payload = pk.payload;
    const std::string &names = pk.names;
    const std::vector<int64_t> &name_off = pk.name_off;
    const int32_t n_pairs = pk.n_pairs, n_clusters = (int32_t)cref.size();
    std::vector<int32_t> pair_group((size_t)n_pairs + 1), &n_groups = pk.n_groups;
    n_groups.assign((size_t)n_clusters + 1, 0);
    std::vector<gcb_group_result> &groups = pk.groups;
    groups.resize((size_t)n_pairs + 1);
    std::vector<uint8_t> &out_payload = pk.out_payload;
    out_payload.resize(payload.size() + 16);
    std::vector<uint64_t> &umi = pk.umi;
    int &umi_words = pk.umi_words;
    if (n_pairs > 0) {
        // 2. UMIs of every read (gcb_extract_umi = BamUtil::getUMI on the GPU), then Pair::setLeft / setRight (pair.cpp:188-216)
        std::vector<uint64_t> read_umi;
        std::vector<uint8_t> status((size_t)2 * n_pairs);
        for (;;) {
            read_umi.assign((size_t)2 * n_pairs * umi_words, 0);
            eng.check(eng.extract_umi(eng.ctx, names.data(), name_off.data(), 2 * n_pairs, prefix.c_str(), umi_words, read_umi.data(), status.data()),
                      "gcb_extract_umi");
            bool too_long = false;
            for (uint8_t s : status) too_long |= s != 0;
            if (!too_long) break;
            if (umi_words == GCB_MAX_UMI_WORDS) die("UMI longer than 64 characters");
            umi_words = GCB_MAX_UMI_WORDS;
        }
        lap("gcb_extract_umi", t0);
        umi.assign((size_t)n_pairs * umi_words, 0);
        for (int32_t p = 0; p < n_pairs; p++) {
            const uint64_t *l = &read_umi[(size_t)(2 * p) * umi_words], *r = l + umi_words;
            bool l_empty = true;
            for (int k = 0; k < umi_words; k++) l_empty &= l[k] == 0;
            const bool has_right = reads[(size_t)2 * p + 1].l_qseq >= 0;
            if (has_right && !l_empty && memcmp(l, r, 8 * (size_t)umi_words) != 0) {
                fprintf(stderr, "Mismatched UMI of a pair of reads\n");
                die("The UMI of a read pair should be identical");
            }
            memcpy(&umi[(size_t)p * umi_words], has_right ? r : l, 8 * (size_t)umi_words);
        }
        // 3. Cluster::clusterByUMI over all of them
        gcb_batch b;
        memset(&b, 0, sizeof b);
        b.n_clusters = n_clusters;
        b.n_pairs = n_pairs;
        b.umi_words = umi_words;
        b.cluster_pair_off = cpo.data();
        b.cluster_ref = cref.data();
        b.cluster_flags = cflags.data();
        b.umi = umi.data();
        b.reads = reads.data();
        if (cigar.empty()) cigar.push_back(0);
        b.cigar = cigar.data();
        b.n_cigar_ops = (int64_t)cigar.size();
        b.payload = payload.data();
        b.payload_bytes = (int64_t)payload.size();
        for (int32_t c = 0; c < n_clusters; c++) {
            const int64_t s0 = reads[(size_t)2 * cpo[(size_t)c]].data_off;
            const int64_t s1 = c + 1 < n_clusters ? reads[(size_t)2 * cpo[(size_t)c + 1]].data_off : (int64_t)payload.size();
            if (s1 - s0 > b.max_cluster_bytes) b.max_cluster_bytes = (int32_t)std::min<int64_t>(s1 - s0, 0x7FFFFFFF);
        }
        gcb_result res;
        int64_t out_bytes = 0;
        res.pair_group = pair_group.data();
        res.cluster_n_groups = n_groups.data();
        res.groups = groups.data();
        res.out_payload = out_payload.data();
        res.out_capacity = (int64_t)out_payload.size();
        res.out_bytes = &out_bytes;
        eng.check(eng.consensus_batch(eng.ctx, &b, &res), "gcb_consensus_batch");
        lap("gcb_consensus_batch", t0);
    }
    // (the packed batch is not needed again: the replay reads the log, the results and the records)
    std::vector<uint8_t>().swap(payload);
    std::vector<gcb_read_desc>().swap(reads);
    std::string().swap(pk.names);
}

// 4. replay: what the loops around clusterByUMI do with the returned pairs (gencore.cpp:355-360, 409-414)
void Pipeline::replay(Packed &pk) {
    double t0 = now_s();
    std::vector<Event> &log = pk.log;
    const std::vector<int32_t> &cpo = pk.cpo, &n_groups = pk.n_groups;
    const std::vector<gcb_group_result> &groups = pk.groups;
    const std::vector<uint8_t> &out_payload = pk.out_payload;
    const std::vector<uint64_t> &umi = pk.umi;
    const std::vector<const Rec *> &slot_rec = pk.slot_rec;
    const int umi_words = pk.umi_words;
    const int32_t n_pairs = pk.n_pairs;
    int32_t c = 0;
    std::vector<char> consumed((size_t)2 * n_pairs + 1, 0);
    for (Event &e : log) {
        if (e.kind == Event::PASS) {
            output_bam(e.rec, true);
            continue;
        }
        if (e.kind == Event::CLEAR_OUTSET) {
            output_out_set();
            continue;
        }
        for (ClusterJob &j : e.jobs) {
            if (j.passthrough) {
                for (PairRec &p : j.pairs) output_pair(p.left, p.right);
                continue;
            }
            const int32_t p0 = cpo[(size_t)c], p1 = cpo[(size_t)c + 1], G = n_groups[(size_t)c];
            bool has_umi = false;
            for (int32_t p = p0; p < p1 && !has_umi; p++)
                for (int k = 0; k < umi_words; k++) has_umi |= umi[(size_t)p * umi_words + k] != 0;
            // result order (cluster.cpp:119-183): popped from the back when duplex pairing runs, else group order
            const bool from_back = has_umi && !cli.opt.disable_duplex;
            for (int32_t step = 0; step < G; step++) {
                const int32_t g = from_back ? G - 1 - step : step;
                const gcb_group_result &gr = groups[(size_t)p0 + g];
                if (gr.status != GCB_GROUP_SSCS && gr.status != GCB_GROUP_DCS) continue;
                Rec *outrec[2] = {nullptr, nullptr};
                for (int side = 0; side < 2; side++) {
                    const int32_t t = gr.tmpl_read[side];
                    if (t < 0) continue;
                    Rec *r = const_cast<Rec *>(slot_rec[(size_t)t]);
                    consumed[(size_t)t] = 1;
                    // the consensus bases and qualities (the template record is rewritten in place, group.cpp:503-525)
                    const uint8_t *o = &out_payload[(size_t)gr.out_off[side]];
                    memcpy(r->qual.data(), o, (size_t)r->l_seq);
                    memcpy(r->seq.data(), o + GCB_ALIGN4(r->l_seq), (size_t)(r->l_seq + 1) / 2);
                    // NM:C (group.cpp:527-572)
                    const int inc = gr.mismatch_inc[side];
                    if (inc != 0 && inc <= 5) {
                        size_t idx = 0;
                        const uint8_t *nm = aux_find(r->aux, "NM", &idx);
                        if (nm && *nm == 'C') {
                            const int v = (int)r->aux[idx + 1] + inc;
                            if (v >= 0 && v <= 255) r->aux[idx + 1] = (uint8_t)v;
                        }
                    }
                    outrec[side] = r;
                }
                for (int side = 0; side < 2; side++) {  // BamUtil::copyQName (group.cpp:109-123): the donor's name
                    if (outrec[side] && gr.qname_donor[side] >= 0) outrec[side]->qname = slot_rec[(size_t)gr.qname_donor[side]]->qname;
                }
                for (int side = 0; side < 2; side++) {  // Pair::writeSscsDcsTagBam (pair.cpp:54-68)
                    if (!outrec[side]) continue;
                    std::vector<uint8_t> &aux = outrec[side]->aux;
                    const unsigned fr = (unsigned)std::min(gr.merge_reads, 65535);
                    aux.push_back('F'); aux.push_back('R'); aux.push_back('C'); aux.push_back((uint8_t)fr);
                    if (gr.status == GCB_GROUP_DCS) {
                        const unsigned rr = (unsigned)std::min(gr.reverse_merge_reads, 65535);
                        aux.push_back('R'); aux.push_back('R'); aux.push_back('C'); aux.push_back((uint8_t)rr);
                    }
                }
                output_pair(outrec[0], outrec[1]);
            }
            c++;
        }
        if (e.set_watermark) {
            processed_tid = e.wm_tid;
            processed_pos = e.wm_pos;
        }
    }
    // the records that did not become a consensus were deleted with their Pair
    for (size_t s = 0; s < slot_rec.size(); s++)
        if (slot_rec[s] && !consumed[s]) delete slot_rec[s];
    lap("replay + write", t0);
}

void Pipeline::run() {
    // the engine (library load, CUDA context, reference upload) starts while the first reads are parsed
    engine_ready = std::async(std::launch::async, [this] {
        double t0 = now_s();
        eng.open(cli.engine);
        eng.check(eng.create(&cli.opt, cli.device, &eng.ctx), "gcb_create");
        genome = load_fasta(cli.ref, eng);
        if (!genome.len.empty())
            eng.check(eng.set_reference(eng.ctx, genome.packed.data(), (int64_t)genome.packed.size(), genome.off.data(), genome.len.data(),
                                        (int32_t)genome.len.size()),
                      "gcb_set_reference");
        lap("engine start-up (async)", t0);
    });
    unsigned hw = std::thread::hardware_concurrency();
    WorkerPool pool((int)std::max(2u, std::min(hw ? hw - 1 : 4u, 16u)));
    BgzfReader in;
    in.pool = &pool;
    out.pool = &pool;
    in.fp = cli.input == "-" ? stdin : fopen(cli.input.c_str(), "rb");
    if (!in.fp) die("failed to open " + cli.input);
    out.fp = cli.output == "-" ? stdout : fopen(cli.output.c_str(), "wb");
    if (!out.fp) die("failed to open output " + cli.output);
    hdr = read_header(in);
    if (hdr.names.empty()) die("this SAM file has no header");
    write_header(out, hdr);
    for (int32_t l : hdr.lens) {
        lin_off.push_back(lin_total);
        lin_total += std::max<int64_t>(l, 1);
    }
    prefix = cli.umi_prefix;
    bool first = true;
    int last_tid = -1, last_pos = -1;
    // pairs per engine call; GCB_BATCH_PAIRS: smaller batches for tests of the hand-over between the three threads
    const size_t batch_pairs = getenv("GCB_BATCH_PAIRS") ? (size_t)std::max(1L, atol(getenv("GCB_BATCH_PAIRS"))) : (size_t)200000;
    to_engine.cap = 6;  // (reader and packer do not stop while the CUDA context starts up: 1 to 4 s on the B200 boxes; six batches ~ 2.3 GB of reads and payload)
    std::thread pack_thread([this] { pack_stage(); }), engine_thread([this] { engine_stage(); }), writer_thread([this] { writer_stage(); });
    double t_read = now_s();
    Rec *b = new Rec();
    while (read_record(in, *b)) {
        if (first) {  // gencore.cpp:207-221
            if (prefix == "auto") prefix = b->qname.find("umi_") != std::string::npos ? "umi" : b->qname.find("UMI_") != std::string::npos ? "UMI" : "";
            first = false;
        }
        if (b->tid < last_tid || (b->tid == last_tid && b->pos < last_pos)) {
            if (b->tid >= 0 && b->pos >= 0) die("the input is unsorted. Please sort the input first.");
        }
        last_tid = b->tid;
        last_pos = b->pos;
        if (b->tid < 0 || b->pos < 0) {  // unmapped reads are dropped; the first one ends the proper clusters (gencore.cpp:255-266)
            if (!out_set_cleared) {
                if (!clusters_finished) {
                    clusters_finished = true;
                    finish_consensus();
                }
                Event e;
                e.kind = Event::CLEAR_OUTSET;
                log.push_back(std::move(e));
                out_set_cleared = true;
            }
            continue;
        }
        if (b->flag & (0x100 | 0x800)) continue;  // secondary / supplementary (bamutil.cpp:368-373)
        if ((size_t)b->tid >= hdr.lens.size() || b->mtid >= (int32_t)hdr.lens.size()) die("corrupt BAM record");  // (contig ids index the header)
        b->serial = serial++;
        add_to_proper_cluster(b);
        b = new Rec();
        if (pending_pairs >= batch_pairs) {
            lap("read + key", t_read);
            submit_log();
            t_read = now_s();
        }
    }
    lap("read + key", t_read);
    delete b;
    if (!clusters_finished) {
        clusters_finished = true;
        finish_consensus();
    }
    Event e;  // ~Gencore: outputOutSet
    e.kind = Event::CLEAR_OUTSET;
    log.push_back(std::move(e));
    submit_log();
    to_pack.close();
    pack_thread.join();
    engine_thread.join();
    writer_thread.join();
    double t_close = now_s();
    out.close();
    lap("flush output", t_close);
    if (in.fp != stdin) fclose(in.fp);
    // every output is written and closed: leave without gcb_destroy and without the teardown of the CUDA context and of the
    // runtime's exit handlers (measured on the B200 boxes: gcb_destroy 0.01 to 1.8 s, the exit handlers 0.3 to 1.3 s);
    // GCB_FAST_EXIT=0: the ordinary way out
    const char *fe = getenv("GCB_FAST_EXIT");
    if (!fe || strcmp(fe, "0") != 0) {
        double t_end = g_t_start;
        lap("main, before exit", t_end);
        fflush(nullptr);
        _exit(0);
    }
    eng.destroy(eng.ctx);
    lap("gcb_destroy", t_close);
}

// --merge: the outputs of the N processes of one sharded run joined into one BAM in the order of the reference's output set
// (gencore.h:19-47; records tied on all five fields keep shard order).  Every shard's output is in that order already.
void merge_shards(const Cli &cli) {
    unsigned hw = std::thread::hardware_concurrency();
    WorkerPool pool((int)std::max(2u, std::min(hw ? hw - 1 : 4u, 16u)));
    const size_t n = cli.merge_inputs.size();
    std::vector<std::unique_ptr<BgzfReader>> in(n);
    std::vector<Rec> head(n);
    std::vector<bool> live(n, false);
    BgzfWriter out;
    out.pool = &pool;
    out.fp = cli.output == "-" ? stdout : fopen(cli.output.c_str(), "wb");
    if (!out.fp) die("failed to open output " + cli.output);
    for (size_t k = 0; k < n; k++) {
        in[k].reset(new BgzfReader());
        in[k]->pool = &pool;
        in[k]->fp = fopen(cli.merge_inputs[k].c_str(), "rb");
        if (!in[k]->fp) die("failed to open " + cli.merge_inputs[k]);
        BamHeader h = read_header(*in[k]);
        if (k == 0) write_header(out, h);
        live[k] = read_record(*in[k], head[k]);
        head[k].serial = k;  // ties: the lower shard first
    }
    OutComp less;
    for (;;) {
        int best = -1;
        for (size_t k = 0; k < n; k++)
            if (live[k] && (best < 0 || less(&head[k], &head[(size_t)best]))) best = (int)k;
        if (best < 0) break;
        write_record(out, head[(size_t)best]);
        live[(size_t)best] = read_record(*in[(size_t)best], head[(size_t)best]);
        head[(size_t)best].serial = (uint64_t)best;
    }
    out.close();
    for (size_t k = 0; k < n; k++) fclose(in[k]->fp);
}

}  // namespace

int main(int argc, char **argv) {
    Pipeline p;
    Cli &c = p.cli;
    // the defaults of Options::Options (options.cpp:4-40) without needing the engine library yet
    memset(&c.opt, 0, sizeof c.opt);
    c.opt.duplex_mismatch_threshold = 2; c.opt.cluster_size_req = 1; c.opt.base_score_req = 6;
    c.opt.high_quality = 30; c.opt.moderate_quality = 20; c.opt.low_quality = 15;
    c.opt.score_high = 8; c.opt.score_moderate = 6; c.opt.score_low = 4; c.opt.score_bad = 2;
    c.opt.skip_low_complexity_cluster_threshold = 1000; c.opt.score_percent_req = 0.8;
    {   // the engine library lies next to the binary's directory: $GENCORE_B200_ENGINE, else relative to /proc/self/exe (argv[0] is
        // only a name when the binary was found through PATH)
        const char *env = getenv("GENCORE_B200_ENGINE");
        char exe[4096];
        const ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
        std::string self = n > 0 ? std::string(exe, (size_t)n) : std::string(argv[0]);
        const size_t slash = self.rfind('/');
        c.engine = env ? std::string(env) : (slash == std::string::npos ? std::string(".") : self.substr(0, slash)) + "/../csrc/libgencore_b200.so";
    }
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char *what) -> std::string {
            if (i + 1 >= argc) die(std::string("option needs value: ") + what);
            return argv[++i];
        };
        if (a == "-i" || a == "--in") c.input = need("in");
        else if (a == "-o" || a == "--out") c.output = need("out");
        else if (a == "-r" || a == "--ref") c.ref = need("ref");
        else if (a == "-u" || a == "--umi_prefix") c.umi_prefix = need("umi_prefix");
        else if (a == "-s" || a == "--supporting_reads") c.opt.cluster_size_req = atoi(need("supporting_reads").c_str());
        else if (a == "-a" || a == "--ratio_threshold") c.opt.score_percent_req = atof(need("ratio_threshold").c_str());
        else if (a == "-c" || a == "--score_threshold") c.opt.base_score_req = atoi(need("score_threshold").c_str());
        else if (a == "-d" || a == "--umi_diff_threshold") c.umi_diff_threshold = atoi(need("umi_diff_threshold").c_str());
        else if (a == "-D" || a == "--duplex_diff_threshold") c.opt.duplex_mismatch_threshold = atoi(need("duplex_diff_threshold").c_str());
        else if (a == "--high_qual") c.opt.high_quality = atoi(need("high_qual").c_str());
        else if (a == "--moderate_qual") c.opt.moderate_quality = atoi(need("moderate_qual").c_str());
        else if (a == "--low_qual") c.opt.low_quality = atoi(need("low_qual").c_str());
        else if (a == "-x" || a == "--duplex_only") c.opt.duplex_only = 1;
        else if (a == "--no_duplex") c.opt.disable_duplex = 1;
        else if (a == "--engine") c.engine = need("engine");   // (not a reference flag) the C-ABI library to load
        else if (a == "--device") c.device = atoi(need("device").c_str());
        else if (a == "--shard") {  // (not a reference flag) i/N: this process takes the i-th of N coordinate windows
            const std::string v = need("shard");
            if (sscanf(v.c_str(), "%d/%d", &c.shard_index, &c.shard_count) != 2 || c.shard_count < 1 || c.shard_index < 0 || c.shard_index >= c.shard_count)
                die("--shard wants i/N with 0 <= i < N");
        } else if (a == "--merge") {  // (not a reference flag) every remaining argument is a shard's output; -o names the result
            while (i + 1 < argc) c.merge_inputs.push_back(argv[++i]);
            if (c.merge_inputs.empty()) die("--merge wants the shards' BAM files");
        }
        else if (a == "-b" || a == "--bed" || a == "-j" || a == "--json" || a == "-h" || a == "--html" || a == "--coverage_sampling" ||
                 a == "--quit_after_contig") need(a.c_str());  // reports and debugging aids are not part of this tool
        else if (a == "--debug") {}
        else die("unrecognized option: " + a);
    }
    if (!c.merge_inputs.empty()) {
        merge_shards(c);
        return 0;
    }
    // Options::validate (options.cpp:42-111), the checks that touch these fields
    if (c.ref.empty()) die("need option: --ref");
    if (c.opt.duplex_only && c.opt.disable_duplex) die("You cannot enable both duplex_only and no_duplex");
    if (c.opt.score_percent_req > 1.0 || c.opt.score_percent_req < 0.5) die("ratio_threshold cannot be greater than 1.0 or less than 0.5");
    if (c.opt.cluster_size_req > 10 || c.opt.cluster_size_req < 1) die("supporting_reads cannot be less than 1 or greater than 10");
    if (c.opt.base_score_req > 10 || c.opt.base_score_req < 1) die("score_threshold cannot be less than 1 or greater than 10");
    if (c.opt.high_quality > 40 || c.opt.high_quality < 20) die("high_qual cannot be greater than 40 or less than 20");
    if (c.opt.moderate_quality > 35 || c.opt.moderate_quality < 15) die("moderate_qual cannot be greater than 35 or less than 15");
    if (c.opt.low_quality > 30 || c.opt.low_quality < 8) die("low_qual cannot be greater than 30 or less than 8");
    if (c.umi_diff_threshold < 0 || c.umi_diff_threshold > 10) die("umi_diff_threshold cannot be negative or greater than 10");
    if (c.opt.low_quality > c.opt.moderate_quality) die("low_qual cannot be greater than moderate_qual");
    if (c.opt.moderate_quality > c.opt.high_quality) die("moderate_qual cannot be greater than high_qual");
    if (c.opt.duplex_mismatch_threshold < 0 || c.opt.duplex_mismatch_threshold > 10) die("duplex_diff_threshold cannot be negative or greater than 10");
    p.run();
    double t_end = g_t_start;
    lap("main, before exit", t_end);
    return 0;
}
