// mini-hts: BGZF + BAM record codec + aux helpers, just enough for the reference to link.
// TEST INFRASTRUCTURE ONLY (see htslib/sam.h in this directory).  Written from the SAM/BAM
// specification; SAM *text* I/O is not implemented (no config uses it) — sam_open() on a
// non-BGZF input or a "w" (text) output fails loudly.
#include "htslib/sam.h"
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace {

const size_t BGZF_BLOCK_PAYLOAD = 0xff00;  // htslib fills 0xff00 bytes per block

struct Bgzf {
    FILE *fp = nullptr;
    bool writing = false;
    bool own_fp = true;
    // read side
    std::vector<uint8_t> inbuf;   // inflated bytes of the current block
    size_t inpos = 0;
    bool eof = false;
    // write side
    std::vector<uint8_t> outbuf;  // pending plain bytes
    int level = 6;
};

bool read_exact(FILE *fp, void *dst, size_t n) { return fread(dst, 1, n, fp) == n; }

// Load and inflate the next BGZF member.  Returns 1 ok, 0 clean EOF, -1 error.
int bgzf_next_block(Bgzf *z) {
    uint8_t hdr[12];
    size_t got = fread(hdr, 1, 12, z->fp);
    if (got == 0) { z->eof = true; return 0; }
    if (got != 12 || hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) return -1;
    uint16_t xlen = hdr[10] | (hdr[11] << 8);
    std::vector<uint8_t> extra(xlen);
    if (!read_exact(z->fp, extra.data(), xlen)) return -1;
    int bsize = -1;
    for (size_t i = 0; i + 4 <= extra.size();) {
        uint16_t slen = extra[i + 2] | (extra[i + 3] << 8);
        if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2) bsize = extra[i + 4] | (extra[i + 5] << 8);
        i += 4 + slen;
    }
    if (bsize < 0) return -1;
    long clen = (long)bsize + 1 - 12 - xlen - 8;
    if (clen < 0) return -1;
    std::vector<uint8_t> comp(clen);
    if (clen && !read_exact(z->fp, comp.data(), clen)) return -1;
    uint8_t tail[8];
    if (!read_exact(z->fp, tail, 8)) return -1;
    uint32_t crc = tail[0] | (tail[1] << 8) | (tail[2] << 16) | ((uint32_t)tail[3] << 24);
    uint32_t isize = tail[4] | (tail[5] << 8) | (tail[6] << 16) | ((uint32_t)tail[7] << 24);
    z->inbuf.resize(isize);
    z->inpos = 0;
    if (isize) {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) return -1;
        zs.next_in = comp.data();
        zs.avail_in = (uInt)clen;
        zs.next_out = z->inbuf.data();
        zs.avail_out = isize;
        int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) return -1;
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), z->inbuf.data(), isize) != crc) return -1;
    }
    return 1;
}

// Returns bytes read (< n only at EOF), -1 on error.
long bgzf_read(Bgzf *z, void *dst, size_t n) {
    uint8_t *out = (uint8_t *)dst;
    size_t done = 0;
    while (done < n) {
        if (z->inpos == z->inbuf.size()) {
            if (z->eof) break;
            int rc = bgzf_next_block(z);
            if (rc < 0) return -1;
            if (rc == 0) break;
            continue;
        }
        size_t take = z->inbuf.size() - z->inpos;
        if (take > n - done) take = n - done;
        memcpy(out + done, z->inbuf.data() + z->inpos, take);
        z->inpos += take;
        done += take;
    }
    return (long)done;
}

int bgzf_write_block(Bgzf *z, const uint8_t *src, size_t n) {
    uint8_t comp[70000];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, z->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    zs.next_in = (Bytef *)src;
    zs.avail_in = (uInt)n;
    zs.next_out = comp;
    zs.avail_out = sizeof comp;
    int rc = deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    size_t total = 18 + clen + 8;
    uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
                       (uint8_t)((total - 1) & 0xff), (uint8_t)((total - 1) >> 8)};
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n);
    uint8_t tail[8] = {(uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24),
                       (uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    if (fwrite(hdr, 1, 18, z->fp) != 18) return -1;
    if (clen && fwrite(comp, 1, clen, z->fp) != clen) return -1;
    if (fwrite(tail, 1, 8, z->fp) != 8) return -1;
    return 0;
}

int bgzf_flush(Bgzf *z, bool all) {
    size_t off = 0;
    while (z->outbuf.size() - off >= BGZF_BLOCK_PAYLOAD || (all && off < z->outbuf.size())) {
        size_t n = z->outbuf.size() - off;
        if (n > BGZF_BLOCK_PAYLOAD) n = BGZF_BLOCK_PAYLOAD;
        if (bgzf_write_block(z, z->outbuf.data() + off, n) < 0) return -1;
        off += n;
    }
    z->outbuf.erase(z->outbuf.begin(), z->outbuf.begin() + off);
    return 0;
}

int bgzf_write(Bgzf *z, const void *src, size_t n) {
    const uint8_t *p = (const uint8_t *)src;
    z->outbuf.insert(z->outbuf.end(), p, p + n);
    if (z->outbuf.size() >= BGZF_BLOCK_PAYLOAD) return bgzf_flush(z, false);
    return 0;
}

inline uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void put32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back(x & 0xff); v.push_back((x >> 8) & 0xff); v.push_back((x >> 16) & 0xff); v.push_back(x >> 24);
}
inline void put16(std::vector<uint8_t> &v, uint16_t x) { v.push_back(x & 0xff); v.push_back(x >> 8); }

// SAMv1 §5.3 reg2bin
int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

void ensure_data(bam1_t *b, size_t need) {
    if (need > b->m_data) {
        size_t m = need + (need >> 2) + 32;
        b->data = (uint8_t *)realloc(b->data, m);
        b->m_data = (uint32_t)m;
    }
}

int aux_type_size(uint8_t t) {
    switch (t) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'd': return 8;
        default: return 0;
    }
}

// advance past the value whose type byte is at s; returns NULL on malformed data
const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end) {
    if (s >= end) return nullptr;
    uint8_t t = *s++;
    int sz = aux_type_size(t);
    if (sz) return s + sz <= end ? s + sz : nullptr;
    if (t == 'Z' || t == 'H') {
        while (s < end && *s) ++s;
        return s < end ? s + 1 : nullptr;
    }
    if (t == 'B') {
        if (s + 5 > end) return nullptr;
        int esz = aux_type_size(*s);
        uint32_t n = le32(s + 1);
        s += 5 + (size_t)esz * n;
        return (esz && s <= end) ? s : nullptr;
    }
    return nullptr;
}

}  // namespace

struct minihts_file {
    Bgzf z;
};

extern "C" {

samFile *sam_open(const char *fn, const char *mode) {
    bool wr = mode && mode[0] == 'w';
    if (wr && !strchr(mode, 'b')) {
        fprintf(stderr, "[mini-hts] SAM text output is not implemented (%s)\n", fn);
        return nullptr;
    }
    samFile *f = new minihts_file();
    f->z.writing = wr;
    if (strcmp(fn, "-") == 0) {
        f->z.fp = wr ? stdout : stdin;
        f->z.own_fp = false;
    } else {
        f->z.fp = fopen(fn, wr ? "wb" : "rb");
    }
    if (!f->z.fp) { delete f; return nullptr; }
    return f;
}

int sam_close(samFile *fp) {
    if (!fp) return 0;
    int rc = 0;
    if (fp->z.writing) {
        if (bgzf_flush(&fp->z, true) < 0) rc = -1;
        static const uint8_t eofblk[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
                                           0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (fwrite(eofblk, 1, 28, fp->z.fp) != 28) rc = -1;
        if (fflush(fp->z.fp) != 0) rc = -1;
    }
    if (fp->z.own_fp && fclose(fp->z.fp) != 0) rc = -1;
    delete fp;
    return rc;
}

bam_hdr_t *bam_hdr_init(void) { return (bam_hdr_t *)calloc(1, sizeof(bam_hdr_t)); }

void bam_hdr_destroy(bam_hdr_t *h) {
    if (!h) return;
    if (h->target_name) {
        for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
        free(h->target_name);
    }
    free(h->target_len);
    free(h->text);
    free(h);
}

bam_hdr_t *sam_hdr_read(samFile *fp) {
    uint8_t buf[8];
    if (bgzf_read(&fp->z, buf, 4) != 4 || memcmp(buf, "BAM\1", 4) != 0) {
        fprintf(stderr, "[mini-hts] input is not a BAM file (SAM text input is not implemented)\n");
        return nullptr;
    }
    bam_hdr_t *h = bam_hdr_init();
    if (bgzf_read(&fp->z, buf, 4) != 4) { bam_hdr_destroy(h); return nullptr; }
    h->l_text = le32(buf);
    h->text = (char *)malloc(h->l_text + 1);
    if (bgzf_read(&fp->z, h->text, h->l_text) != (long)h->l_text) { bam_hdr_destroy(h); return nullptr; }
    h->text[h->l_text] = 0;
    if (bgzf_read(&fp->z, buf, 4) != 4) { bam_hdr_destroy(h); return nullptr; }
    h->n_targets = (int32_t)le32(buf);
    h->target_name = (char **)calloc(h->n_targets ? h->n_targets : 1, sizeof(char *));
    h->target_len = (uint32_t *)calloc(h->n_targets ? h->n_targets : 1, sizeof(uint32_t));
    for (int i = 0; i < h->n_targets; ++i) {
        if (bgzf_read(&fp->z, buf, 4) != 4) { bam_hdr_destroy(h); return nullptr; }
        uint32_t l = le32(buf);
        h->target_name[i] = (char *)malloc(l + 1);
        if (bgzf_read(&fp->z, h->target_name[i], l) != (long)l) { bam_hdr_destroy(h); return nullptr; }
        h->target_name[i][l] = 0;
        if (bgzf_read(&fp->z, buf, 4) != 4) { bam_hdr_destroy(h); return nullptr; }
        h->target_len[i] = le32(buf);
    }
    return h;
}

int sam_hdr_write(samFile *fp, const bam_hdr_t *h) {
    std::vector<uint8_t> v;
    v.insert(v.end(), {'B', 'A', 'M', 1});
    put32(v, (uint32_t)h->l_text);
    v.insert(v.end(), (const uint8_t *)h->text, (const uint8_t *)h->text + h->l_text);
    put32(v, (uint32_t)h->n_targets);
    for (int i = 0; i < h->n_targets; ++i) {
        uint32_t l = (uint32_t)strlen(h->target_name[i]) + 1;
        put32(v, l);
        v.insert(v.end(), (const uint8_t *)h->target_name[i], (const uint8_t *)h->target_name[i] + l);
        put32(v, h->target_len[i]);
    }
    if (bgzf_write(&fp->z, v.data(), v.size()) < 0) return -1;
    // htslib ends the header on a block boundary
    return bgzf_flush(&fp->z, true);
}

bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}

// >=0 ok, -1 EOF, < -1 error (gencore.cpp:205 loops on >= 0)
int sam_read1(samFile *fp, bam_hdr_t *, bam1_t *b) {
    uint8_t x[36];
    long got = bgzf_read(&fp->z, x, 4);
    if (got == 0) return -1;
    if (got != 4) return -2;
    uint32_t block_len = le32(x);
    if (block_len < 32) return -3;
    if (bgzf_read(&fp->z, x + 4, 32) != 32) return -4;
    bam1_core_t *c = &b->core;
    c->tid = (int32_t)le32(x + 4);
    c->pos = (int32_t)le32(x + 8);
    uint32_t l_read_name = x[12];
    c->qual = x[13];
    c->bin = x[14] | (x[15] << 8);
    c->n_cigar = x[16] | (x[17] << 8);
    c->flag = x[18] | (x[19] << 8);
    c->l_qseq = (int32_t)le32(x + 20);
    c->mtid = (int32_t)le32(x + 24);
    c->mpos = (int32_t)le32(x + 28);
    c->isize = (int32_t)le32(x + 32);
    // htslib pads the in-memory qname with NULs so the CIGAR is 4-byte aligned
    c->l_extranul = (uint8_t)((4 - (l_read_name & 3)) & 3);
    c->l_qname = (uint16_t)(l_read_name + c->l_extranul);
    size_t rest = block_len - 32;
    size_t l_data = rest + c->l_extranul;
    ensure_data(b, l_data);
    b->l_data = (int)l_data;
    if (bgzf_read(&fp->z, b->data, l_read_name) != (long)l_read_name) return -4;
    for (int i = 0; i < c->l_extranul; ++i) b->data[l_read_name + i] = 0;
    if (bgzf_read(&fp->z, b->data + c->l_qname, rest - l_read_name) != (long)(rest - l_read_name)) return -4;
    return (int)block_len + 4;
}

int sam_write1(samFile *fp, const bam_hdr_t *, const bam1_t *b) {
    const bam1_core_t *c = &b->core;
    uint32_t l_read_name = c->l_qname - c->l_extranul;
    uint32_t block_len = 32 + (uint32_t)b->l_data - c->l_extranul;
    std::vector<uint8_t> v;
    v.reserve(block_len + 4);
    put32(v, block_len);
    put32(v, (uint32_t)c->tid);
    put32(v, (uint32_t)(int32_t)c->pos);
    v.push_back((uint8_t)l_read_name);
    v.push_back(c->qual);
    put16(v, c->bin);
    put16(v, (uint16_t)c->n_cigar);
    put16(v, c->flag);
    put32(v, (uint32_t)c->l_qseq);
    put32(v, (uint32_t)c->mtid);
    put32(v, (uint32_t)(int32_t)c->mpos);
    put32(v, (uint32_t)(int32_t)c->isize);
    v.insert(v.end(), b->data, b->data + l_read_name);
    v.insert(v.end(), b->data + c->l_qname, b->data + b->l_data);
    if (bgzf_write(&fp->z, v.data(), v.size()) < 0) return -1;
    return (int)v.size();
}

// pointer to the TYPE byte of the tag's value, or NULL (group.cpp:532-535 depends on this)
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    const uint8_t *s = bam_get_aux(b);
    const uint8_t *end = b->data + b->l_data;
    while (s && s + 3 <= end) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return (uint8_t *)(s + 2);
        s = aux_skip(s + 2, end);
    }
    return nullptr;
}

int64_t bam_aux2i(const uint8_t *s) {
    uint8_t t = *s++;
    switch (t) {
        case 'c': return (int8_t)s[0];
        case 'C': return s[0];
        case 's': return (int16_t)(s[0] | (s[1] << 8));
        case 'S': return (uint16_t)(s[0] | (s[1] << 8));
        case 'i': return (int32_t)le32(s);
        case 'I': return le32(s);
        default: return 0;
    }
}

char *bam_aux2Z(const uint8_t *s) {
    uint8_t t = *s++;
    if (t == 'Z' || t == 'H') return (char *)s;
    return nullptr;
}

int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data) {
    size_t need = (size_t)b->l_data + 3 + len;
    ensure_data(b, need);
    uint8_t *p = b->data + b->l_data;
    p[0] = tag[0];
    p[1] = tag[1];
    p[2] = (uint8_t)type;
    memcpy(p + 3, data, len);
    b->l_data = (int)need;
    return 0;
}

hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar) {
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; ++k)
        if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2) l += bam_cigar_oplen(cigar[k]);
    return l;
}

bam1_t *minihts_make_record(const char *qname, int l_qname_with_nul, int32_t tid, hts_pos_t pos,
                            uint16_t flag, int32_t mtid, hts_pos_t mpos, hts_pos_t isize,
                            const uint32_t *cigar, uint32_t n_cigar,
                            const uint8_t *seq4, const uint8_t *qual, int32_t l_qseq,
                            const uint8_t *aux, int l_aux) {
    bam1_t *b = bam_init1();
    bam1_core_t *c = &b->core;
    c->tid = tid; c->pos = pos; c->flag = flag; c->mtid = mtid; c->mpos = mpos; c->isize = isize;
    c->qual = 60;
    c->n_cigar = n_cigar;
    c->l_qseq = l_qseq;
    c->l_extranul = (uint8_t)((4 - (l_qname_with_nul & 3)) & 3);
    c->l_qname = (uint16_t)(l_qname_with_nul + c->l_extranul);
    size_t l_data = c->l_qname + 4 * (size_t)n_cigar + (size_t)((l_qseq + 1) >> 1) + l_qseq + l_aux;
    ensure_data(b, l_data);
    b->l_data = (int)l_data;
    memset(b->data, 0, c->l_qname);
    memcpy(b->data, qname, l_qname_with_nul - 1);
    if (n_cigar) memcpy(bam_get_cigar(b), cigar, 4 * (size_t)n_cigar);
    memcpy(bam_get_seq(b), seq4, (l_qseq + 1) >> 1);
    memcpy(bam_get_qual(b), qual, l_qseq);
    if (l_aux) memcpy(bam_get_aux(b), aux, l_aux);
    hts_pos_t rlen = n_cigar ? bam_cigar2rlen((int)n_cigar, cigar) : 1;
    c->bin = (uint16_t)reg2bin(pos, pos + (rlen > 0 ? rlen : 1));
    return b;
}

}  // extern "C"
