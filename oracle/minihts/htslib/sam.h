/*
 * mini-hts: the slice of htslib's <htslib/sam.h> that OpenGene/gencore uses.
 *
 * TEST INFRASTRUCTURE ONLY.  htslib is an un-vendored, un-pinned dependency of the
 * reference (Makefile:16 `-lhts`) and is absent from this image, so the reference's
 * own sources (/root/reference/src/*.cpp) are compiled UNCHANGED against this header
 * and oracle/minihts/minihts.cpp to build oracle/_ref/.  Nothing under gencore_b200/
 * includes or links this.
 *
 * Written from the SAM/BAM specification (SAMv1 §4) and the documented htslib >= 1.10
 * public API; no htslib source was available.  Only what the reference touches exists:
 *   types   bam1_core_t, bam1_t, bam_hdr_t, samFile
 *   funcs   sam_open sam_close sam_hdr_read sam_hdr_write sam_read1 sam_write1
 *           bam_init1 bam_destroy1 bam_hdr_destroy bam_aux_get bam_aux2i bam_aux2Z
 *           bam_aux_append bam_cigar2rlen
 *   macros  bam_get_qname/cigar/seq/qual/aux, bam_cigar_op/oplen/opchr, BAM_C*, BAM_F*
 */
#ifndef MINIHTS_SAM_H
#define MINIHTS_SAM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9

#define BAM_CIGAR_STR   "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define BAM_CIGAR_TYPE  0x3C1A7

#define bam_cigar_op(c)    ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_opchr(c) (BAM_CIGAR_STR "??????"[bam_cigar_op(c)])
#define bam_cigar_type(o)  (BAM_CIGAR_TYPE >> ((o) << 1) & 3)

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy;
} bam1_t;

typedef struct sam_hdr_t {
    int32_t n_targets, ignore_sam_err;
    size_t l_text;
    uint32_t *target_len;
    const int8_t *cigar_tab;
    char **target_name;
    char *text;
    void *sdict;
    void *hrecs;
    uint32_t ref_count;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

struct minihts_file;
typedef struct minihts_file samFile;
typedef struct minihts_file htsFile;

#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i)   ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

samFile *sam_open(const char *fn, const char *mode);
int sam_close(samFile *fp);
bam_hdr_t *sam_hdr_read(samFile *fp);
int sam_hdr_write(samFile *fp, const bam_hdr_t *h);
int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b);
int sam_write1(samFile *fp, const bam_hdr_t *h, const bam1_t *b);

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam_hdr_t *bam_hdr_init(void);
void bam_hdr_destroy(bam_hdr_t *h);

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
char *bam_aux2Z(const uint8_t *s);
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar);

/* mini-hts extension used only by the oracle harness: build an in-memory record. */
bam1_t *minihts_make_record(const char *qname, int l_qname_with_nul, int32_t tid, hts_pos_t pos,
                            uint16_t flag, int32_t mtid, hts_pos_t mpos, hts_pos_t isize,
                            const uint32_t *cigar, uint32_t n_cigar,
                            const uint8_t *seq4, const uint8_t *qual, int32_t l_qseq,
                            const uint8_t *aux, int l_aux);

#ifdef __cplusplus
}
#endif
#endif
