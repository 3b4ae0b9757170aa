/*
 * gencore_b200.h — C ABI of libgencore_b200.so, the B200 (sm_100a) consensus engine that
 * replaces the hot path of OpenGene/gencore v0.17.2:
 *
 *     Cluster::clusterByUMI            /root/reference/src/cluster.cpp:55-188   (cluster.h:26)
 *       -> Group::consensusMerge       group.cpp:68-134
 *            -> consensusMergeBam      group.cpp:136-318  (BamUtil::isPartOf bamutil.cpp:204-255)
 *                 -> Pair::computeScore   pair.cpp:88-172
 *                 -> Group::makeConsensus group.cpp:320-579 (BamUtil::getRefOffset bamutil.cpp:293-314,
 *                                                            Reference::getData reference.cpp:33-71)
 *       -> Cluster::duplexMerge(Bam)   cluster.cpp:190-244
 *
 * The reference has no plugin/FFI seam; its in-process seam is
 *     vector<Pair*> Cluster::clusterByUMI(int umiDiffThreshold, Stats*, Stats*, bool crossContig)
 * called once per (tid,left,right) cluster from gencore.cpp:355 and gencore.cpp:409.  This ABI
 * is that call batched over thousands of clusters: the caller packs the clusters it would have
 * handed to clusterByUMI into one gcb_batch, and gets back, for every UMI family ("group"), what
 * the reference would have returned in its Pair (consensus bases/quals, FR/RR counts, which record
 * is the template, what to do with qname and NM).  Plain pointers and sizes only; no global state;
 * never calls exit().  Every function returns GCB_OK (0) or a negative gcb_status.
 *
 * Encoding conventions
 *   bases    BAM 4-bit codes, two per byte, EVEN index in the HIGH nibble (bam_get_seq layout).
 *   quals    raw phred bytes (bam_get_qual layout).
 *   payload  one record per read: l_qseq qual bytes at data_off, then the packed bases at
 *            data_off + GCB_ALIGN4(l_qseq).  Records are laid out in read-slot order; those of one
 *            cluster are contiguous and the cluster's first record starts on a 16-byte boundary,
 *            so cluster c's slab is [reads[2*off[c]].data_off, reads[2*off[c+1]].data_off) (the
 *            last one ends at payload_bytes) and one TMA bulk copy stages a whole cluster.  Every
 *            pair has its side-0 read (Cluster::addRead always fills mLeft first, cluster.cpp:260).
 *   UMI      `umi_words` u64 per pair; character k of the UMI string is the 4-bit field at bits
 *            [60-4*(k%16), 64-4*(k%16)) of word k/16 (first character in the MOST significant
 *            nibble): A=1 C=2 G=3 T=4 _=5, 0 = past the end.  (The reference's getUMI,
 *            bamutil.cpp:23-112, can only produce these five characters.)  With this code umiDiff
 *            (cluster.cpp:41-53) is the number of differing fields, and comparing the words as
 *            unsigned integers, word 0 first, is std::string order (the map<string,int> order of
 *            cluster.cpp:57-76).
 *   genome   the reference's own packing (fastareader.cpp:139-152): A=1 T=2 C=3 G=4 other=0,
 *            EVEN index in the LOW nibble, one byte string per contig.
 */
#ifndef GENCORE_B200_H
#define GENCORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCB_ABI_VERSION 2
#define GCB_ALIGN4(x) (((x) + 3) & ~3)
#define GCB_MAX_UMI_WORDS 4 /* UMIs up to 64 characters */

typedef enum gcb_status {
    GCB_OK = 0,
    GCB_ERR_ARG = -1,          /* NULL / inconsistent argument */
    GCB_ERR_CUDA = -2,         /* CUDA runtime failure; see gcb_last_error */
    GCB_ERR_NO_DEVICE = -3,    /* no sm_100 device: there is NO CPU fallback */
    GCB_ERR_CAPACITY = -4,     /* out_payload too small */
    GCB_ERR_MALFORMED = -5     /* batch violates a documented precondition */
} gcb_status;

/* The fields of Options (options.h:15-61, defaults options.cpp:4-40) that the hot path reads. */
typedef struct gcb_options {
    int32_t duplex_mismatch_threshold; /* -D  duplexMismatchThreshold   (2)  cluster.cpp:137 */
    int32_t cluster_size_req;          /* -s  clusterSizeReq            (1)  cluster.cpp:138,159,174 */
    int32_t base_score_req;            /* -c  baseScoreReq              (6)  group.cpp:423,466 */
    int32_t high_quality;              /*     highQuality               (30) */
    int32_t moderate_quality;          /*     moderateQuality           (20) */
    int32_t low_quality;               /*     lowQuality                (15) */
    int32_t score_high;                /* scoreOfNotOverlappedHighQual     (8) pair.cpp:77-86 */
    int32_t score_moderate;            /* scoreOfNotOverlappedModerateQual (6) */
    int32_t score_low;                 /* scoreOfNotOverlappedLowQual      (4) */
    int32_t score_bad;                 /* scoreOfNotOverlappedBadQual      (2) */
    int32_t skip_low_complexity_cluster_threshold; /* (1000) group.cpp:142,231 */
    int32_t duplex_only;               /* -x  cluster.cpp:159,174 */
    int32_t disable_duplex;            /* --no_duplex cluster.cpp:119 */
    int32_t reserved;
    double score_percent_req;          /* -a  scorePercentReq (0.8) group.cpp:462 */
} gcb_options;

/* One per read slot; slot = 2*pair + side, side 0 = Pair::mLeft, 1 = Pair::mRight (pair.h:50-51). */
typedef struct gcb_read_desc {
    int64_t data_off;   /* byte offset of this read's record in payload (multiple of 4) */
    int32_t l_qseq;     /* core.l_qseq; < 0 marks an empty slot (pair without mRight) */
    int32_t pos;        /* core.pos */
    int32_t isize;      /* core.isize (group.cpp:363 only tests != 0) */
    int32_t cigar_off;  /* index of the first op in gcb_batch.cigar */
    uint16_t n_cigar;   /* core.n_cigar */
    uint16_t l_qname;   /* core.l_qname INCLUDING htslib's NUL padding to a multiple of 4
                           (bamutil.cpp:19-21, group.cpp:90-96,115-122) */
    uint32_t reserved;
} gcb_read_desc;

#define GCB_CLUSTER_CROSS_CONTIG 0x01u      /* crossContig argument (gencore.cpp:355: right < 0) */
#define GCB_CLUSTER_UMI_THR_SHIFT 4         /* umiDiffThreshold argument in bits 4..7 (0..10) */

typedef struct gcb_batch {
    int32_t n_clusters;
    int32_t n_pairs;
    int32_t umi_words;               /* 1..GCB_MAX_UMI_WORDS */
    int32_t max_cluster_bytes;       /* hint: largest cluster slab in payload bytes, 0 = unknown */
    const int32_t *cluster_pair_off; /* [n_clusters+1]; pairs of cluster c are [off[c], off[c+1]) in the
                                        iteration order of Cluster::mPairs (map<string qname+pad>, cluster.h:45) */
    const int32_t *cluster_ref;      /* [n_clusters] genome contig index for the cluster's tid, or -1 when
                                        the BAM contig name is not in the FASTA (reference.cpp:51-58) */
    const uint8_t *cluster_flags;    /* [n_clusters] GCB_CLUSTER_* */
    const uint64_t *umi;             /* [n_pairs*umi_words] Pair::mUMI */
    const gcb_read_desc *reads;      /* [2*n_pairs] */
    const uint32_t *cigar;           /* BAM cigar ops (len<<4|op), indexed by cigar_off */
    int64_t n_cigar_ops;
    const uint8_t *payload;
    int64_t payload_bytes;           /* multiple of 16 */
} gcb_batch;

/* gcb_group_result.status */
#define GCB_GROUP_DROPPED 0        /* below -s, or SSCS under -x (cluster.cpp:159-165,174-182) */
#define GCB_GROUP_SSCS 1           /* kept, FR tag only */
#define GCB_GROUP_DCS 2            /* kept, duplex: FR + RR tags (cluster.cpp:137-146) */
#define GCB_GROUP_DUPLEX_PARTNER 3 /* was p2 of a duplex merge: consumed (cluster.cpp:151-152) */
#define GCB_GROUP_DUPLEX_DIFF 4    /* p1 dropped: diff > -D (cluster.cpp:147-150) */
#define GCB_GROUP_DUPLEX_SMALL 5   /* p1 dropped: FR+RR < -s (cluster.cpp:144-146) */

/* One per group slot; slot = cluster_pair_off[c] + g, g = creation order in cluster.cpp:66-100. */
typedef struct gcb_group_result {
    int64_t out_off[2];      /* consensus record of side s in out_payload (same record layout as payload) */
    int32_t tmpl_read[2];    /* read slot that became the consensus (group.cpp:270-281), -1 = NULL */
    int32_t qname_donor[2];  /* read slot whose qname BamUtil::copyQName copies over side s's
                                (group.cpp:110-123), -1 = keep */
    int32_t diff[2];         /* makeConsensus return value (group.cpp:579) */
    int32_t mismatch_inc[2]; /* group.cpp:521-525; != 0 and <= 5: caller patches NM:C (group.cpp:568-572);
                                > 5: bases+quals were rolled back (group.cpp:538-566) */
    int32_t merge_reads;     /* Pair::mMergeReads  -> FR:C (pair.cpp:54-60) */
    int32_t reverse_merge_reads; /* Pair::mReverseMergeReads -> RR:C, 0 unless status == DCS */
    int32_t status;          /* GCB_GROUP_* */
    int32_t duplex_partner;  /* group index (within the cluster) merged with, -1 = none */
    int32_t duplex_diff;     /* duplexMerge return (cluster.cpp:190-198) when duplex_partner >= 0 */
    int32_t umi_pair;        /* pair index whose UMI the consensus Pair carries, -1 = empty UMI */
} gcb_group_result;

typedef struct gcb_result {
    int32_t *pair_group;       /* [n_pairs] group index of each pair inside its cluster */
    int32_t *cluster_n_groups; /* [n_clusters] */
    gcb_group_result *groups;  /* [n_pairs]; entries g >= cluster_n_groups[c] are zero-filled */
    uint8_t *out_payload;      /* consensus records, cluster order, group order, side order */
    int64_t out_capacity;      /* bytes available at out_payload (payload_bytes is always enough) */
    int64_t *out_bytes;        /* [1] bytes used */
} gcb_result;

/* stage mask for gcb_consensus_batch_device: the four stages of the path */
#define GCB_STAGE_UMI_GROUP 0x1u       /* cluster.cpp:55-100  -> pair_group, cluster_n_groups */
#define GCB_STAGE_SELECT_TEMPLATE 0x2u /* group.cpp:136-313   -> tmpl_read, out_off, qname_donor, umi_pair */
#define GCB_STAGE_SCORE_VOTE 0x4u      /* pair.cpp:88-172 + group.cpp:320-579 -> out_payload, diff, mismatch_inc */
#define GCB_STAGE_DUPLEX 0x8u          /* cluster.cpp:102-244 -> status, FR/RR, duplex merge of out_payload */
#define GCB_STAGE_ALL 0xFu
/* measurement only: the parts of GCB_STAGE_SCORE_VOTE one at a time — per-tile preparation (tile_prep2_kernel), then the
 * vote, or the vote's two halves: the ring kernel (vote_ring_kernel), then slow columns + rollback + the generic kernel's tiles */
#define GCB_STAGE_VOTE_PREP_ONLY 0x10u
#define GCB_STAGE_VOTE_ONLY 0x20u
#define GCB_STAGE_VOTE_FAST_ONLY 0x40u
#define GCB_STAGE_VOTE_REST_ONLY 0x80u

typedef struct gcb_ctx gcb_ctx;

int gcb_abi_version(void);
void gcb_default_options(gcb_options *opt);

/* Binds one context to one GPU (one per process/stream). Fails with GCB_ERR_NO_DEVICE when the
 * device is absent or not sm_100: the engine has no CPU path. */
int gcb_create(const gcb_options *opt, int device, gcb_ctx **out);
void gcb_destroy(gcb_ctx *ctx);
const char *gcb_last_error(const gcb_ctx *ctx);

/* Reference genome (Reference::getData, reference.cpp:33-71).  packed4/contig_* are HOST pointers;
 * contig_off[i] = byte offset of contig i in packed4, contig_len[i] = its length in bases. */
int gcb_set_reference(gcb_ctx *ctx, const uint8_t *packed4, int64_t packed_bytes,
                      const int64_t *contig_off, const int64_t *contig_len, int32_t n_contigs);
/* Same, but packed4 is already a DEVICE pointer (e.g. filled by an NCCL broadcast); the context
 * borrows it — the caller keeps it alive.  contig_* are host pointers. */
int gcb_set_reference_device(gcb_ctx *ctx, const uint8_t *packed4_dev, int64_t packed_bytes,
                             const int64_t *contig_off, const int64_t *contig_len, int32_t n_contigs);

/* The call that replaces the loop over Cluster::clusterByUMI: HOST buffers in, HOST buffers out
 * (pinned memory recommended); copies, the four kernels and the copy back run on the context's
 * stream and the call returns when the results are in `result`. */
int gcb_consensus_batch(gcb_ctx *ctx, const gcb_batch *batch, gcb_result *result);

/* Same path with every pointer inside *batch and *result a DEVICE pointer (the structs themselves
 * live on the host).  Asynchronous on `stream` (a cudaStream_t; NULL = the context's stream);
 * `stages` selects which kernels run (GCB_STAGE_*), later stages read what earlier ones left in
 * *result and in the context's workspace. */
int gcb_consensus_batch_device(gcb_ctx *ctx, const gcb_batch *batch, gcb_result *result,
                               uint32_t stages, void *stream);

/* Device-side error flag raised by the last batch (GCB_OK, GCB_ERR_CAPACITY, GCB_ERR_MALFORMED);
 * synchronises the stream. */
int gcb_batch_status(gcb_ctx *ctx, void *stream);

/* BamUtil::getUMI(string qname, const string& prefix) (bamutil.cpp:40-112) for n strings at once, encoded as the
 * `umi` field of gcb_batch wants it.  names = the strings back to back (no terminators needed), name_off[n+1] their
 * byte offsets; the caller passes the MI:Z tag value instead of the qname when the record has one (bamutil.cpp:23-38).
 * prefix = Options::umiPrefix ("" = no-prefix mode), at most 31 characters.  status[i] = 0, or 1 when the UMI has more
 * than 16*umi_words characters (its code is truncated).  HOST buffers; runs on the context's stream and returns when
 * out_umi / status are filled. */
int gcb_extract_umi(gcb_ctx *ctx, const char *names, const int64_t *name_off, int32_t n, const char *prefix, int32_t umi_words,
                    uint64_t *out_umi, uint8_t *status);

/* FastaReader::readAll + to4bits (fastareader.cpp:58-152) on the device: the text of a FASTA file in, the genome in the layout
 * gcb_set_reference wants out.  text / outputs are HOST buffers.  Contig i (file order) has contig_len[i] bases at byte
 * contig_off[i] (16-byte aligned) of packed4_out; its id — the header up to the first space, fastareader.cpp:98-102 — is
 * text[name_off[i] .. name_off[i] + name_len[i]).  (The reference keeps its contigs in a map by id: of two contigs with the same
 * id the LATER one wins.)  Everything the reference's reader does with odd input is reproduced — text before the first '>',
 * lower case, CR, digits, '-' and '*', blank lines that swallow the next line — except the two header shapes it
 * does not read as headers ('>' directly before a line end or another '>'): GCB_ERR_MALFORMED.  GCB_ERR_CAPACITY when there
 * are more than max_contigs contigs or packed_cap (n / 2 + 16 * max_contigs + 16 is always enough) is too small. */
int gcb_pack_fasta(gcb_ctx *ctx, const char *text, int64_t n, int32_t max_contigs, uint8_t *packed4_out, int64_t packed_cap,
                   int64_t *contig_off, int64_t *contig_len, int64_t *name_off, int32_t *name_len, int32_t *n_contigs, int64_t *packed_bytes);

/* Tuning knob of gcb_consensus_batch: payload bytes per pipeline chunk (default 48 MiB; at most 16 chunks per call,
 * chunks are whole clusters).  Results do not depend on it. */
int gcb_set_chunk_bytes(gcb_ctx *ctx, int64_t bytes);

/* Page-locked host memory for the arrays of a gcb_batch / gcb_result handed to gcb_consensus_batch (pageable memory works,
 * at a fraction of the link rate).  `write_combined` memory is for buffers the host only WRITES (a packed batch): the copy
 * engine reads it without snooping the CPU caches; reading it back on the host is very slow (on the round-1 B200 box both
 * kinds upload at the same 44 GB/s).  NULL on failure. */
void *gcb_host_alloc(size_t bytes, int write_combined);
void gcb_host_free(void *p);

/* Tuning / test aid; results never depend on it.  key 2 = force the vote's tile window (14 or 15 = log2 bytes, 0 = automatic),
 * key 3 = lanes per cluster in umi_group_kernel / select_template_kernel (8, 16, 32; 0 = by mean cluster size),
 * key 5 = non-zero: every tile is voted by the generic kernel (score_vote_kernel) instead of the ring kernel,
 * key 6 = non-zero: gcb_consensus_batch prints its timeline (enqueue, copy-in stream, results) on stderr. */
int gcb_set_debug(gcb_ctx *ctx, int key, int value);

/* Tuning knob: bytes of the slow-column queue between vote_ring_kernel and slow_columns_kernel (0 = sized from the payload).
 * Tiles whose slow columns do not fit are voted by the generic kernel; results do not depend on it. */
int gcb_set_slow_queue_bytes(gcb_ctx *ctx, int64_t bytes);

/* What Cluster::clusterByUMI adds to preStats and postStats (cluster.cpp:102,136,143,157,161,172,176,184-186 calling
 * Stats::addCluster / addMolecule / addSSCS / addDCS, stats.cpp:122-141), summed on the device over every cluster this context
 * has processed with GCB_STAGE_DUPLEX since the last reset.  The counters are additive: ranks sum them with one all-reduce. */
#define GCB_MAX_SUPPORTING_READS 100 /* stats.h:15 */
typedef struct gcb_cluster_stats {
    int64_t pre_cluster, pre_multi_cluster;                              /* preStats->addCluster */
    int64_t pre_molecule, pre_molecule_se, pre_molecule_pe, pre_uncounted; /* preStats->addMolecule; uncounted: >= 100 supporting reads */
    int64_t post_cluster, post_multi_cluster, post_sscs, post_dcs;       /* postStats->addCluster / addSSCS / addDCS */
    int64_t pre_hist[GCB_MAX_SUPPORTING_READS];                          /* Stats::mSupportingHistgram */
} gcb_cluster_stats;
/* Waits for the context's stream, copies the counters to *out and, with reset != 0, zeroes them. */
int gcb_get_cluster_stats(gcb_ctx *ctx, gcb_cluster_stats *out, int reset);

/* Stats::statDepth (stats.cpp:56-83, called by Stats::addRead for every mapped read) over n reads given as host arrays of
 * core.tid / core.pos / core.l_qseq: ADDS every read's bases to depth[bin_off[tid] + k], k-th bin of `coverage_step` bases of
 * contig tid (Options::coverageStep, 10000 by default), where bin_off[t] = sum over earlier contigs of 1 + target_len / step
 * (Stats::makeGenomeDepthBuf, stats.cpp:40-46).  `depth` is a host array of bin_off[n_targets] counters (zero them for a
 * fresh Stats object; they are additive over batches and over ranks).  The rules of the reference hold bit for bit: reads with
 * tid outside [0, n_targets) or ending in or beyond the contig's last bin are not counted. */
int gcb_stat_depth(gcb_ctx *ctx, const int32_t *tid, const int32_t *pos, const int32_t *l_qseq, int64_t n, int32_t coverage_step,
                   const int64_t *target_len, int32_t n_targets, int64_t *depth);

/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t gcb_launch_count(const gcb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GENCORE_B200_H */
