// gcbbridge.h — the binding a gencore maintainer adds to call libgencore_b200.so instead of Cluster::clusterByUMI
// (INTEGRATION.md).  It is compiled INTO the reference (integration/patch_reference.py splices it into the two call sites,
// gencore.cpp:355 and gencore.cpp:409; oracle/Makefile target `bridge` builds oracle/_ref/gencore_bridged from the
// reference's own sources), so it speaks the reference's types: Cluster, Pair, bam1_t, Stats, Reference.
//
//   add(cluster, thr, crossContig)   instead of   cluster->clusterByUMI(thr, pre, post, crossContig) + the outputPair loop:
//                                    appends the cluster's pairs (map order) to the pending gcb_batch and keeps the cluster.
//   flush(pre, post, outputPair)     one gcb_consensus_batch for everything pending, then — cluster by cluster, in the order
//                                    they were added — exactly what clusterByUMI would have returned and done: the consensus
//                                    Pair objects (template records rewritten in place, qname copied, NM patched, FR / RR
//                                    tags written by the reference's own Pair::writeSscsDcsTag), the Stats side effects, the
//                                    other records freed.
//
// No consensus arithmetic here: packing (what the header's "Encoding conventions" ask for) and replay (what
// gcb_group_result says).  The engine is loaded with dlopen ($GENCORE_B200_ENGINE, default libgencore_b200.so on the
// loader path), so the reference needs no CUDA toolchain to build.  Fail-stop like the reference (util.h error_exit).
#ifndef GCB_BRIDGE_H
#define GCB_BRIDGE_H

#include <dlfcn.h>
#include <string.h>

#include <functional>
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
#include "util.h"

class GcbBridge {
public:
    GcbBridge(Options *opt, bam_hdr_t *hdr) : mOptions(opt), mCtx(NULL), mUmiWords(1) {
        const char *path = getenv("GENCORE_B200_ENGINE");
        mLib = dlopen(path && *path ? path : "libgencore_b200.so", RTLD_NOW | RTLD_LOCAL);
        if (!mLib) error_exit(std::string("gencore_b200: cannot load the engine library: ") + dlerror());
        load(fDefaultOptions, "gcb_default_options");
        load(fCreate, "gcb_create");
        load(fDestroy, "gcb_destroy");
        load(fLastError, "gcb_last_error");
        load(fSetReference, "gcb_set_reference");
        load(fConsensusBatch, "gcb_consensus_batch");
        gcb_options o;
        fDefaultOptions(&o);  // options.h:15-61 -> gcb_options
        o.duplex_mismatch_threshold = opt->duplexMismatchThreshold;
        o.cluster_size_req = opt->clusterSizeReq;
        o.base_score_req = opt->baseScoreReq;
        o.high_quality = opt->highQuality;
        o.moderate_quality = opt->moderateQuality;
        o.low_quality = opt->lowQuality;
        o.score_high = opt->scoreOfNotOverlappedHighQual;
        o.score_moderate = opt->scoreOfNotOverlappedModerateQual;
        o.score_low = opt->scoreOfNotOverlappedLowQual;
        o.score_bad = opt->scoreOfNotOverlappedBadQual;
        o.skip_low_complexity_cluster_threshold = opt->skipLowComplexityClusterThreshold;
        o.duplex_only = opt->duplexOnly ? 1 : 0;
        o.disable_duplex = opt->disableDuplex ? 1 : 0;
        o.score_percent_req = opt->scorePercentReq;
        if (fCreate(&o, 0, &mCtx) != GCB_OK) error_exit("gencore_b200: no sm_100 device (the engine has no CPU path)");
        packGenome(hdr);
    }
    ~GcbBridge() {
        if (mCtx) fDestroy(mCtx);
    }

    size_t pendingPairs() const { return mUmi.size(); }

    // Replaces the call of clusterByUMI: the cluster's pairs join the pending batch.  The bridge owns `c` from here on.
    void add(Cluster *c, int umiDiffThreshold, bool crossContig) {
        int tid = -1;
        mPairOff.push_back((int32_t)mPairs.size());
        while (mPayload.size() & 15) mPayload.push_back(0);  // a cluster's slab starts on a 16-byte boundary
        std::map<std::string, Pair *>::iterator it;
        for (it = c->mPairs.begin(); it != c->mPairs.end(); it++) {  // map order: the order clusterByUMI walks (cluster.h:45)
            Pair *p = it->second;
            mPairs.push_back(p);
            mUmi.push_back(p->getUMI());
            bam1_t *side[2] = {p->mLeft, p->mRight};
            for (int s = 0; s < 2; s++) {
                gcb_read_desc d;
                memset(&d, 0, sizeof d);
                d.l_qseq = -1;
                bam1_t *b = side[s];
                if (b) {
                    if (tid < 0) tid = b->core.tid;
                    d.data_off = (int64_t)mPayload.size();
                    d.l_qseq = b->core.l_qseq;
                    d.pos = b->core.pos;
                    d.isize = b->core.isize;
                    d.cigar_off = (int32_t)mCigar.size();
                    d.n_cigar = (uint16_t)b->core.n_cigar;
                    d.l_qname = (uint16_t)b->core.l_qname;  // htslib's padded length (bamutil.cpp:19-21)
                    const uint32_t *cig = bam_get_cigar(b);
                    mCigar.insert(mCigar.end(), cig, cig + b->core.n_cigar);
                    const int l = b->core.l_qseq, qb = GCB_ALIGN4(l), sb = GCB_ALIGN4((l + 1) / 2);
                    const size_t at = mPayload.size();
                    mPayload.resize(at + (size_t)qb + (size_t)sb, 0);
                    memcpy(&mPayload[at], bam_get_qual(b), (size_t)l);
                    memcpy(&mPayload[at + (size_t)qb], bam_get_seq(b), (size_t)(l + 1) / 2);
                }
                mReads.push_back(d);
                mRecs.push_back(b);
            }
        }
        int ref = -1;
        if (tid >= 0 && tid < (int)mTidToContig.size()) ref = mTidToContig[tid];
        mClusterRef.push_back(ref);
        mClusterFlags.push_back((uint8_t)((crossContig ? GCB_CLUSTER_CROSS_CONTIG : 0) | (umiDiffThreshold << GCB_CLUSTER_UMI_THR_SHIFT)));
        mClusters.push_back(c);
    }

    // One engine call for everything pending, then the replay of clusterByUMI's return values and side effects.
    void flush(Stats *preStats, Stats *postStats, const std::function<void(Pair *)> &outputPair) {
        const int32_t nc = (int32_t)mClusters.size(), np = (int32_t)mPairs.size();
        if (nc == 0) return;
        mPairOff.push_back(np);
        while (mPayload.size() & 15) mPayload.push_back(0);
        // UMIs as 4-bit codes, first character in the most significant nibble (the header's encoding conventions)
        size_t longest = 0;
        for (size_t i = 0; i < mUmi.size(); i++) longest = std::max(longest, mUmi[i].size());
        mUmiWords = (int)std::max<size_t>(1, (longest + 15) / 16);
        if (mUmiWords > GCB_MAX_UMI_WORDS) error_exit("gencore_b200: UMI longer than 64 characters");
        std::vector<uint64_t> umi((size_t)np * mUmiWords, 0);
        for (int32_t p = 0; p < np; p++)
            for (size_t k = 0; k < mUmi[p].size(); k++) {
                const char ch = mUmi[p][k];
                const uint64_t code = ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : ch == 'T' ? 4 : ch == '_' ? 5 : 0;
                if (code == 0) error_exit("gencore_b200: a UMI character that BamUtil::getUMI cannot produce");
                umi[(size_t)p * mUmiWords + k / 16] |= code << (60 - 4 * (k % 16));
            }
        if (mCigar.empty()) mCigar.push_back(0);
        gcb_batch b;
        memset(&b, 0, sizeof b);
        b.n_clusters = nc;
        b.n_pairs = np;
        b.umi_words = mUmiWords;
        b.cluster_pair_off = mPairOff.data();
        b.cluster_ref = mClusterRef.data();
        b.cluster_flags = mClusterFlags.data();
        b.umi = umi.data();
        b.reads = mReads.data();
        b.cigar = mCigar.data();
        b.n_cigar_ops = (int64_t)mCigar.size();
        b.payload = mPayload.data();
        b.payload_bytes = (int64_t)mPayload.size();
        for (int32_t c = 0; c < nc; c++) {
            const int64_t s0 = mReads[(size_t)2 * mPairOff[c]].data_off;
            const int64_t s1 = c + 1 < nc ? mReads[(size_t)2 * mPairOff[c + 1]].data_off : b.payload_bytes;
            if (s1 - s0 > b.max_cluster_bytes) b.max_cluster_bytes = (int32_t)std::min<int64_t>(s1 - s0, 0x7FFFFFFF);
        }
        std::vector<int32_t> pairGroup((size_t)np), nGroups((size_t)nc);
        std::vector<gcb_group_result> groups((size_t)np);
        std::vector<uint8_t> out(mPayload.size() + 16);
        int64_t outBytes = 0;
        gcb_result r;
        r.pair_group = pairGroup.data();
        r.cluster_n_groups = nGroups.data();
        r.groups = groups.data();
        r.out_payload = out.data();
        r.out_capacity = (int64_t)out.size();
        r.out_bytes = &outBytes;
        if (fConsensusBatch(mCtx, &b, &r) != GCB_OK) error_exit(std::string("gencore_b200: ") + fLastError(mCtx));

        std::vector<char> kept(mRecs.size(), 0);  // records that live on in a returned Pair
        for (int32_t c = 0; c < nc; c++) {
            const int32_t p0 = mPairOff[c], p1 = mPairOff[c + 1], G = nGroups[c];
            bool hasUMI = false;
            for (int32_t p = p0; p < p1; p++) hasUMI = hasUMI || !mUmi[p].empty();
            preStats->addCluster(G > 1);  // cluster.cpp:102
            // result order (cluster.cpp:119-183): popped from the back when duplex pairing runs, else group order
            const bool fromBack = hasUMI && !mOptions->disableDuplex;
            int returned = 0;
            std::vector<Pair *> result;
            for (int32_t step = 0; step < G; step++) {
                const int32_t g = fromBack ? G - 1 - step : step;
                const gcb_group_result &gr = groups[(size_t)p0 + g];
                if (gr.status == GCB_GROUP_DUPLEX_PARTNER) continue;  // consumed by its partner's turn
                const bool PE = gr.tmpl_read[0] >= 0 && gr.tmpl_read[1] >= 0;
                int supporting = gr.merge_reads;
                if (gr.duplex_partner >= 0) supporting += groups[(size_t)p0 + gr.duplex_partner].merge_reads;
                preStats->addMolecule(supporting, PE);  // cluster.cpp:136,157,172
                if (gr.status != GCB_GROUP_SSCS && gr.status != GCB_GROUP_DCS) continue;  // the reference deleted this Pair
                Pair *p = new Pair(mOptions);
                bam1_t *rec[2] = {NULL, NULL};
                for (int s = 0; s < 2; s++) {
                    const int32_t t = gr.tmpl_read[s];
                    if (t < 0) continue;
                    bam1_t *o = mRecs[(size_t)t];
                    kept[(size_t)t] = 1;
                    // the consensus bases and qualities: the template record rewritten in place (group.cpp:503-525)
                    const uint8_t *src = &out[(size_t)gr.out_off[s]];
                    memcpy(bam_get_qual(o), src, (size_t)o->core.l_qseq);
                    memcpy(bam_get_seq(o), src + GCB_ALIGN4(o->core.l_qseq), (size_t)(o->core.l_qseq + 1) / 2);
                    const int inc = gr.mismatch_inc[s];  // NM:C (group.cpp:527-572)
                    if (inc != 0 && inc <= 5) {
                        const char tagNM[2] = {'N', 'M'};
                        uint8_t *nm = (uint8_t *)bam_aux_get(o, tagNM);
                        if (nm && *nm == 'C') {
                            const int v = (int)nm[1] + inc;
                            if (v >= 0 && v <= 255) nm[1] = (uint8_t)v;
                        }
                    }
                    rec[s] = o;
                }
                // BamUtil::copyQName (group.cpp:109-123): donors keep their own names until every copy is made
                std::string donorName[2];
                bam1_t *donor[2] = {NULL, NULL};
                for (int s = 0; s < 2; s++)
                    if (rec[s] && gr.qname_donor[s] >= 0) donor[s] = mRecs[(size_t)gr.qname_donor[s]];
                for (int s = 0; s < 2; s++)
                    if (donor[s] && donor[s] != rec[s]) BamUtil::copyQName(donor[s], rec[s]);
                if (rec[0]) p->setLeft(rec[0]);
                if (rec[1]) p->setRight(rec[1]);
                p->mMergeLeftDiff = gr.diff[0];
                p->mMergeRightDiff = gr.diff[1];
                p->mMergeReads = gr.merge_reads;
                if (gr.status == GCB_GROUP_DCS) {
                    p->setDuplex(gr.reverse_merge_reads);  // cluster.cpp:140
                    postStats->addDCS();
                } else {
                    postStats->addSSCS();
                }
                p->writeSscsDcsTag();  // pair.cpp:43-68, the reference's own
                result.push_back(p);
                returned++;
            }
            if (returned > 0) postStats->addCluster(returned > 1);  // cluster.cpp:184-186
            for (size_t i = 0; i < result.size(); i++) {  // gencore.cpp:356-360 / 410-414
                outputPair(result[i]);
                delete result[i];
            }
        }
        // everything else dies with its Pair and its Cluster, as in the reference (cluster.cpp:147-152,163-165,178-182)
        for (size_t i = 0; i < mPairs.size(); i++) {
            Pair *p = mPairs[i];
            if (kept[2 * i]) p->mLeft = NULL;
            if (kept[2 * i + 1]) p->mRight = NULL;
            delete p;
        }
        for (size_t i = 0; i < mClusters.size(); i++) {
            mClusters[i]->mPairs.clear();
            delete mClusters[i];
        }
        mClusters.clear(); mPairs.clear(); mRecs.clear(); mUmi.clear(); mReads.clear(); mCigar.clear(); mPayload.clear();
        mPairOff.clear(); mClusterRef.clear(); mClusterFlags.clear();
    }

private:
    template <typename F>
    void load(F &f, const char *name) {
        f = (F)dlsym(mLib, name);
        if (!f) error_exit(std::string("gencore_b200: the engine library lacks ") + name);
    }
    // Reference::getData's store (reference.cpp:33-71) is already packed as gcb_set_reference wants it
    // (fastareader.cpp:139-152): the contigs the BAM header names, in tid order, each on a 16-byte boundary.
    void packGenome(bam_hdr_t *hdr) {
        mTidToContig.assign(hdr ? (size_t)hdr->n_targets : 0, -1);
        Reference *ref = Reference::instance(mOptions);
        if (!hdr || !ref || !ref->mRef) return;
        std::vector<uint8_t> packed;
        std::vector<int64_t> off, len;
        for (int t = 0; t < hdr->n_targets; t++) {
            const std::string name(hdr->target_name[t]);
            if (ref->mRef->mAllContigs.count(name) == 0) continue;
            const long n = ref->mRef->mAllContigSizes[name];
            mTidToContig[(size_t)t] = (int)off.size();
            off.push_back((int64_t)packed.size());
            len.push_back((int64_t)n);
            const unsigned char *data = ref->mRef->mAllContigs[name];
            packed.insert(packed.end(), data, data + (n + 1) / 2);
            while (packed.size() & 15) packed.push_back(0);
        }
        if (off.empty()) return;
        if (fSetReference(mCtx, packed.data(), (int64_t)packed.size(), off.data(), len.data(), (int32_t)off.size()) != GCB_OK)
            error_exit(std::string("gencore_b200: ") + fLastError(mCtx));
    }

    Options *mOptions;
    void *mLib;
    gcb_ctx *mCtx;
    int mUmiWords;
    void (*fDefaultOptions)(gcb_options *);
    int (*fCreate)(const gcb_options *, int, gcb_ctx **);
    void (*fDestroy)(gcb_ctx *);
    const char *(*fLastError)(const gcb_ctx *);
    int (*fSetReference)(gcb_ctx *, const uint8_t *, int64_t, const int64_t *, const int64_t *, int32_t);
    int (*fConsensusBatch)(gcb_ctx *, const gcb_batch *, gcb_result *);
    std::vector<int> mTidToContig;
    // the pending batch
    std::vector<Cluster *> mClusters;
    std::vector<Pair *> mPairs;
    std::vector<bam1_t *> mRecs;  // [2 * pair + side]
    std::vector<std::string> mUmi;
    std::vector<gcb_read_desc> mReads;
    std::vector<uint32_t> mCigar;
    std::vector<uint8_t> mPayload;
    std::vector<int32_t> mPairOff, mClusterRef;
    std::vector<uint8_t> mClusterFlags;
};

#endif
