/*
 * gencore_oracle.h — CPU restatement of OpenGene/gencore's consensus hot path on the packed
 * batch format of include/gencore_b200.h.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the CHECKER; never linked, imported or executed by
 * anything under gencore_b200/ (the product has no CPU path).
 *
 * Parity status: PINNED.  (i) the reference's own known-answer vectors for this path
 * (BamUtil::test bamutil.cpp:385-423, Cluster::test cluster.cpp:275-288) are checked in
 * tests/test_oracle_kat.py; (ii) the restatement is compared bit-for-bit against the reference
 * itself (oracle/_ref/libgencore_ref.so = /root/reference/src compiled unchanged + ref_harness.cpp)
 * on randomised batches in tests/test_oracle_vs_reference.py, and against committed golden
 * vectors generated from it (tests/golden/, tests/make_golden.py).
 */
#ifndef GENCORE_ORACLE_H
#define GENCORE_ORACLE_H

#include "gencore_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gco_genome {
    const uint8_t *packed4;    /* fastareader.cpp:139-152 packing */
    const int64_t *contig_off; /* byte offset of each contig in packed4 */
    const int64_t *contig_len; /* bases */
    int32_t n_contigs;
} gco_genome;

/* clusterByUMI over every cluster of the batch; host pointers everywhere. */
int gco_consensus_batch(const gcb_options *opt, const gco_genome *genome, const gcb_batch *batch,
                        gcb_result *result);

/* string-level helpers restated for the reference's known-answer tests */
int gco_umi_diff(const char *umi1, const char *umi2);           /* cluster.cpp:41-53 */
int gco_is_duplex(const char *umi1, const char *umi2);          /* cluster.cpp:246-258 */
int gco_get_umi(const char *qname, const char *prefix, char *out, int cap); /* bamutil.cpp:40-112 */
int gco_encode_umi(const char *umi, uint64_t *words, int n_words);          /* string -> 4-bit fields */
void gco_pack_genome(const char *bases, int64_t n, uint8_t *packed4);       /* fastareader.cpp:139-152 */

#ifdef __cplusplus
}
#endif
#endif
