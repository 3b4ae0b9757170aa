// ref_harness.cpp — drives the UNMODIFIED reference classes (Cluster/Group/Pair/BamUtil/Reference,
// compiled from /root/reference/src by oracle/Makefile) over a packed gcb_batch.
// TEST INFRASTRUCTURE ONLY: builds into oracle/_ref/libgencore_ref.so, loaded only by tests/ and
// bench.py's reference arm.  Contains no consensus arithmetic of its own: it rebuilds the bam1_t
// records and Cluster objects gencore.cpp:295-316 would have built, calls
// Cluster::clusterByUMI (cluster.cpp:55) exactly as gencore.cpp:355/409 do, and reports what the
// returned Pair objects contain.
#include <time.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "bamutil.h"
#include "cluster.h"
#include "gencore_b200.h"
#include "options.h"
#include "pair.h"
#include "reference.h"
#include "stats.h"

extern "C" {

typedef struct gcr_pair_result {
    int32_t cluster;
    int32_t slot[2];       // read slot of Pair::mLeft / mRight (the template), -1 = NULL
    int32_t name_slot[2];  // a read slot of the cluster whose ORIGINAL qname equals the returned record's qname
    int32_t l_qname[2];    // core.l_qname of the returned record
    int32_t diff[2];       // mMergeLeftDiff / mMergeRightDiff
    int32_t nm[2];         // NM value of the returned record (-1: no tag)
    int32_t fr[2];         // FR:C tag value read back (-1: absent)
    int32_t rr[2];         // RR:C tag value read back (-1: absent)
    int32_t merge_reads, reverse_merge_reads, is_duplex;
    int32_t reserved;
    int64_t out_off[2];    // record (quals, then bases at +ALIGN4(l)) in the out buffer, -1 = none
} gcr_pair_result;

typedef struct gcr_stats {
    int64_t pre_cluster, pre_multi_cluster, pre_molecule, pre_molecule_se, pre_molecule_pe, pre_uncounted;
    int64_t pre_hist[MAX_SUPPORTING_READS];
    int64_t post_cluster, post_multi_cluster, post_sscs, post_dcs;
} gcr_stats;

struct gcr_ctx {
    Options opt;
    bam_hdr_t hdr;
    std::vector<std::string> names;
    std::vector<char *> name_ptrs;
    std::vector<uint32_t> lens;
    bool has_ref;
};

// fasta_path may be NULL/"" (no reference).  Contig i of the FASTA must be target i; one extra target
// "__absent__" is appended so that cluster_ref == -1 maps to a contig the FASTA lacks.
gcr_ctx *gcr_create(const gcb_options *o, const char *umi_prefix, const char *fasta_path, int n_targets,
                    const char **target_names, const int64_t *target_lens) {
    gcr_ctx *c = new gcr_ctx();
    Options &opt = c->opt;
    opt.umiPrefix = umi_prefix ? umi_prefix : "";
    opt.duplexMismatchThreshold = o->duplex_mismatch_threshold;
    opt.clusterSizeReq = o->cluster_size_req;
    opt.baseScoreReq = o->base_score_req;
    opt.highQuality = o->high_quality;
    opt.moderateQuality = o->moderate_quality;
    opt.lowQuality = o->low_quality;
    opt.scoreOfNotOverlappedHighQual = (char)o->score_high;
    opt.scoreOfNotOverlappedModerateQual = (char)o->score_moderate;
    opt.scoreOfNotOverlappedLowQual = (char)o->score_low;
    opt.scoreOfNotOverlappedBadQual = (char)o->score_bad;
    opt.skipLowComplexityClusterThreshold = o->skip_low_complexity_cluster_threshold;
    opt.duplexOnly = o->duplex_only != 0;
    opt.disableDuplex = o->disable_duplex != 0;
    opt.scorePercentReq = o->score_percent_req;
    for (int i = 0; i < n_targets; i++) {
        c->names.push_back(target_names[i]);
        c->lens.push_back((uint32_t)target_lens[i]);
    }
    c->names.push_back("__absent__");
    c->lens.push_back(1u << 30);
    for (size_t i = 0; i < c->names.size(); i++) c->name_ptrs.push_back(const_cast<char *>(c->names[i].c_str()));
    memset(&c->hdr, 0, sizeof c->hdr);
    c->hdr.n_targets = (int32_t)c->names.size();
    c->hdr.target_name = c->name_ptrs.data();
    c->hdr.target_len = c->lens.data();
    opt.bamHeader = &c->hdr;
    c->has_ref = fasta_path && fasta_path[0];
    if (c->has_ref) {
        opt.refFile = fasta_path;
        // Reference is a process-wide singleton (reference.cpp:4-11): drop any previous instance
        if (Reference::mInstance) delete Reference::mInstance;
        Reference::instance(&opt);
    } else {
        opt.refFile = "";
        if (Reference::mInstance) delete Reference::mInstance;
        Reference::instance(&opt);
    }
    // the singleton keeps a pointer to opt (reference.cpp:14); keep it pointing at ours
    Reference::mInstance->mOptions = &c->opt;
    return c;
}

void gcr_destroy(gcr_ctx *c) {
    if (Reference::mInstance) delete Reference::mInstance;
    delete c;
}

static int aux_c(bam1_t *b, const char *tag) {
    uint8_t *p = bam_aux_get(b, tag);
    if (!p) return -1;
    return (int)bam_aux2i(p);
}

// qnames: n_pairs NUL-terminated names back to back at qname_off[p]; nm: [2*n_pairs] NM:C value per read.
// Returns the number of Pairs the reference returned (<= max_results written) or < 0 on error.
// *seconds receives the wall time spent inside the Cluster::clusterByUMI calls alone.
int64_t gcr_consensus_batch(gcr_ctx *c, const gcb_batch *b, const char *qnames, const int64_t *qname_off,
                            const uint8_t *nm, gcr_pair_result *results, int64_t max_results, uint8_t *out,
                            int64_t out_capacity, int64_t *out_bytes, gcr_stats *stats, double *seconds) {
    Stats pre(&c->opt), post(&c->opt);
    pre.setPostStats(false);
    post.setPostStats(true);
    int64_t nres = 0, cursor = 0;
    double total = 0;
    const int absent_tid = (int)c->names.size() - 1;
    for (int cl = 0; cl < b->n_clusters; cl++) {
        int p0 = b->cluster_pair_off[cl], p1 = b->cluster_pair_off[cl + 1];
        int tid = b->cluster_ref[cl] >= 0 ? b->cluster_ref[cl] : absent_tid;
        Cluster *cluster = new Cluster(&c->opt);
        std::map<bam1_t *, int> slot_of;
        std::map<std::string, int> slot_of_name;
        for (int p = p0; p < p1; p++) {
            const char *qn = qnames + qname_off[p];
            for (int s = 0; s < 2; s++) {
                const gcb_read_desc &r = b->reads[2 * p + s];
                if (r.l_qseq < 0) continue;
                uint8_t aux[4] = {'N', 'M', 'C', nm ? nm[2 * p + s] : (uint8_t)0};
                bam1_t *rec = minihts_make_record(qn, (int)strlen(qn) + 1, tid, r.pos, 0, tid, 0, r.isize,
                                                  b->cigar + r.cigar_off, r.n_cigar,
                                                  b->payload + r.data_off + GCB_ALIGN4(r.l_qseq),
                                                  b->payload + r.data_off, r.l_qseq, aux, 4);
                if (rec->core.l_qname != r.l_qname) {
                    fprintf(stderr, "gcr: l_qname mismatch for pair %d: batch says %d, record has %d\n", p,
                            (int)r.l_qname, (int)rec->core.l_qname);
                    return -1;
                }
                slot_of[rec] = 2 * p + s;
                if (!slot_of_name.count(qn)) slot_of_name[qn] = 2 * p + s;
                cluster->addRead(rec);  // cluster.cpp:260-273: first qname -> setLeft, second -> setRight
            }
        }
        // the batch must list pairs in the iteration order of Cluster::mPairs
        {
            int p = p0;
            for (std::map<std::string, Pair *>::iterator it = cluster->mPairs.begin(); it != cluster->mPairs.end(); ++it, ++p) {
                bam1_t *l = it->second->mLeft;
                if (!l || slot_of[l] != 2 * p) {
                    fprintf(stderr, "gcr: cluster %d: pairs are not in map<qname> order at pair %d\n", cl, p);
                    return -2;
                }
            }
        }
        int thr = b->cluster_flags[cl] >> GCB_CLUSTER_UMI_THR_SHIFT;
        bool cross = (b->cluster_flags[cl] & GCB_CLUSTER_CROSS_CONTIG) != 0;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        std::vector<Pair *> cs = cluster->clusterByUMI(thr, &pre, &post, cross);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        for (size_t i = 0; i < cs.size(); i++) {
            Pair *p = cs[i];
            if (results && nres < max_results) {
                gcr_pair_result &r = results[nres];
                memset(&r, 0, sizeof r);
                r.cluster = cl;
                r.merge_reads = p->mMergeReads;
                r.reverse_merge_reads = p->mReverseMergeReads;
                r.is_duplex = p->mIsDuplex ? 1 : 0;
                r.diff[0] = p->mMergeLeftDiff;
                r.diff[1] = p->mMergeRightDiff;
                bam1_t *recs[2] = {p->mLeft, p->mRight};
                for (int s = 0; s < 2; s++) {
                    bam1_t *rec = recs[s];
                    r.slot[s] = r.name_slot[s] = -1;
                    r.out_off[s] = -1;
                    r.nm[s] = r.fr[s] = r.rr[s] = -1;
                    if (!rec) continue;
                    r.slot[s] = slot_of.count(rec) ? slot_of[rec] : -2;
                    std::string name(bam_get_qname(rec));
                    r.name_slot[s] = slot_of_name.count(name) ? slot_of_name[name] : -2;
                    r.l_qname[s] = rec->core.l_qname;
                    r.nm[s] = aux_c(rec, "NM");
                    r.fr[s] = aux_c(rec, "FR");
                    r.rr[s] = aux_c(rec, "RR");
                    int l = rec->core.l_qseq;
                    int64_t sz = GCB_ALIGN4(l) + GCB_ALIGN4((l + 1) / 2);
                    if (out && cursor + sz <= out_capacity) {
                        memset(out + cursor, 0, sz);
                        memcpy(out + cursor, bam_get_qual(rec), l);
                        memcpy(out + cursor + GCB_ALIGN4(l), bam_get_seq(rec), (l + 1) / 2);
                        r.out_off[s] = cursor;
                        cursor += sz;
                    }
                }
            }
            nres++;
            delete p;
        }
        delete cluster;
    }
    if (out_bytes) *out_bytes = cursor;
    if (seconds) *seconds = total;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->pre_cluster = pre.mCluster;
        stats->pre_multi_cluster = pre.mMultiMoleculeCluster;
        stats->pre_molecule = pre.mMolecule;
        stats->pre_molecule_se = pre.mMoleculeSE;
        stats->pre_molecule_pe = pre.mMoleculePE;
        stats->pre_uncounted = pre.uncountedSupportingReads;
        for (int i = 0; i < MAX_SUPPORTING_READS; i++) stats->pre_hist[i] = pre.mSupportingHistgram[i];
        stats->post_cluster = post.mCluster;
        stats->post_multi_cluster = post.mMultiMoleculeCluster;
        stats->post_sscs = post.mSSCSNum;
        stats->post_dcs = post.mDCSNum;
    }
    return nres;
}

// the reference's own string-level functions, for the KAT cross-check
int gcr_umi_diff(const char *a, const char *b) { return Cluster::umiDiff(a, b); }
int gcr_is_duplex(const char *a, const char *b) { return Cluster::isDuplex(a, b) ? 1 : 0; }
int gcr_get_umi(const char *qname, const char *prefix, char *out, int cap) {
    std::string u = BamUtil::getUMI(std::string(qname), std::string(prefix));
    if ((int)u.size() + 1 > cap) return -1;
    memcpy(out, u.c_str(), u.size() + 1);
    return (int)u.size();
}
int gcr_self_test(void) { return (BamUtil::test() && Cluster::test()) ? 1 : 0; }

// The reference's own FastaReader (fastareader.cpp) over a file, contig by contig in FILE order (readNext is driven as
// readAll drives it, fastareader.cpp:159-170, without the map that would hide a repeated id): id (NUL-terminated, up to 255
// bytes each), size in bases, and the 4-bit data back to back at 16-byte aligned offsets.  Returns the number of contigs, or
// -1 when something does not fit.
int gcr_fasta_load(const char *path, int max_contigs, char *ids, int64_t *sizes, int64_t *offs, uint8_t *packed, int64_t packed_cap) {
    Options opt;
    int n = 0;
    int64_t at = 0;
    try {
        FastaReader reader(&opt, path);
        while (reader.hasNext()) {
            reader.readNext();
            if (n >= max_contigs) return -1;
            const std::string id = reader.currentID();
            const int64_t bytes = (reader.mCurrentSize + 1) / 2;
            if (id.size() > 255 || at + bytes > packed_cap) return -1;
            memcpy(ids + 256 * n, id.c_str(), id.size() + 1);
            sizes[n] = reader.mCurrentSize;
            offs[n] = at;
            if (bytes) memcpy(packed + at, reader.mCurrentSequence, (size_t)bytes);
            delete[] reader.mCurrentSequence;
            at += (bytes + 15) & ~(int64_t)15;
            n++;
        }
    } catch (...) {
        return -1;
    }
    return n;
}

}  // extern "C"
