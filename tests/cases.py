"""Named parity cases shared by the oracle-vs-reference tests, the golden generator and the GPU tests."""
from __future__ import annotations

import numpy as np

from gencore_b200 import synth
from gencore_b200.abi import Options
from gencore_b200.synth import BatchBuilder, SynthCluster, SynthPair, SynthRead


def _q(n, v=37):
    return np.full(n, v, np.uint8)


def edge_batch(seed: int = 7):
    """Hand-built clusters for the quirks listed in SURVEY §8(a) Q1-Q16."""
    rng = np.random.Generator(np.random.PCG64(seed))
    contigs, genome = synth.random_genome(rng, [5000, 3000])
    ref = contigs[0]
    bb = BatchBuilder(2, "UMI")

    def rd(pos, l, isize, cigar=None, seq=None, qual=None, nm=0):
        s = bytes(ref[pos:pos + l]) if seq is None else seq
        return SynthRead(pos, cigar or f"{l}M", s, _q(l) if qual is None else qual, isize, nm)

    def mut(seq, at, to=None):
        a = bytearray(seq)
        for k in (at if isinstance(at, (list, tuple)) else [at]):
            a[k] = ord(to) if to else ord("ACGT"[("ACGT".index(chr(a[k])) + 1) % 4])
        return bytes(a)

    # 1. singleton pair with mate: computeScore rewrites overlap-mismatch quals, -1 scores (Q1,Q4)
    L = 100
    l = rd(100, L, 150)
    r = rd(150, L, -150)
    r.seq = mut(r.seq, [0, 10, 20])       # overlap mismatches vs left
    r.qual = r.qual.copy(); r.qual[0] = 30; r.qual[10] = 37; r.qual[20] = 5
    l.qual = l.qual.copy(); l.qual[50] = 37; l.qual[60] = 30; l.qual[70] = 37
    bb.add(SynthCluster(0, [SynthPair(b"e1:UMI_AAAAAAAA", "AAAAAAAA", l, r)]))
    # 2. singleton without mate: untouched (group.cpp:73-77)
    bb.add(SynthCluster(0, [SynthPair(b"e2:UMI_CCCCCCCC", "CCCCCCCC", rd(300, L, 0))]))
    # 3. two reads disagreeing, low quals -> reference arbitration + NM changes (Q7,Q8)
    prs = []
    for i in range(2):
        l = rd(400, L, 160); r = rd(460, L, -160)
        if i == 0:
            l.seq = mut(l.seq, [5, 6, 7, 8, 9, 10, 11]); l.qual = l.qual.copy(); l.qual[5:12] = 11
        prs.append(SynthPair(b"e3_%d:UMI_GGGGGGGG" % i, "GGGGGGGG", l, r))
    bb.add(SynthCluster(0, prs))
    # 4. same but the cluster's contig is absent from the FASTA (no refdata)
    prs = []
    for i in range(3):
        l = rd(400, L, 160); r = rd(460, L, -160)
        if i < 2:
            l.seq = mut(l.seq, [5 + i]); l.qual = l.qual.copy(); l.qual[5 + i] = 2
        prs.append(SynthPair(b"e4_%d:UMI_GGGGGGGG" % i, "GGGGGGGG", l, r))
    bb.add(SynthCluster(-1, prs))
    # 5. the family carries a true variant at 7 columns where the (shortest-read) template has the reference base:
    #    the vote writes 7 new mismatches -> mismatchInc 7 > 5 -> rollback (Q12)
    prs = []
    for i in range(5):
        l = rd(700, L, 170); r = rd(770, L, -170)
        l.seq = mut(l.seq, [3, 13, 23, 33, 43, 53, 63]); l.nm = 7
        prs.append(SynthPair(b"e5_%d:UMI_TTTTTTTT" % i, "TTTTTTTT", l, r))
    prs[0].left = rd(700, L - 1, 170)
    bb.add(SynthCluster(0, prs))
    # 6. as 5 with 4 variant columns -> NM patched by +4; and the mirror image (template wrong, family right) -> -4 (Q11)
    prs = []
    for i in range(5):
        l = rd(900, L, 170)
        l.seq = mut(l.seq, [3, 13, 23, 33]); l.nm = 4
        prs.append(SynthPair(b"e6_%d:UMI_TTTTTTTA" % i, "TTTTTTTA", l, rd(970, L, -170)))
    prs[0].left = rd(900, L - 1, 170)
    bb.add(SynthCluster(0, prs))
    prs = []
    for i in range(5):
        prs.append(SynthPair(b"e6b_%d:UMI_TTTTTTAA" % i, "TTTTTTAA", rd(1000, L, 170), rd(1070, L, -170)))
    short = rd(1000, L - 1, 170)
    short.seq = mut(short.seq, [3, 13, 23, 33]); short.qual = short.qual.copy(); short.qual[[3, 13, 23, 33]] = 2
    short.nm = 4
    prs[0].left = short
    bb.add(SynthCluster(0, prs))
    # 7. duplex with mismatches incl. the byte-shortcut parity quirk (Q15): mismatches at bases 0 and 4, and 1 and 7
    A, B = "ACGTACGT", "TTGCAAGC"
    top = rd(1200, L, 180); bot = rd(1200, L, 180)
    bot.seq = mut(bot.seq, [0, 4])
    topr = rd(1280, L, -180); botr = rd(1280, L, -180)
    botr.seq = mut(botr.seq, [1, 7])
    bb.add(SynthCluster(0, [SynthPair(b"e7_a:UMI_%s_%s" % (A.encode(), B.encode()), A + "_" + B, top, topr),
                            SynthPair(b"e7_b:UMI_%s_%s" % (B.encode(), A.encode()), B + "_" + A, bot, botr)]))
    # 8. duplex with too many differences -> dropped (cluster.cpp:147-150)
    top = rd(1500, L, 180); bot = rd(1500, L, 180)
    bot.seq = mut(bot.seq, [2, 9, 17, 33])
    bb.add(SynthCluster(0, [SynthPair(b"e8_a:UMI_%s_%s" % (A.encode(), B.encode()), A + "_" + B, top, rd(1580, L, -180)),
                            SynthPair(b"e8_b:UMI_%s_%s" % (B.encode(), A.encode()), B + "_" + A, bot, rd(1580, L, -180))]))
    # 9. cross-contig cluster: left reads only, name donor = shortest then smallest name (group.cpp:80-99)
    prs = [SynthPair(b"e9_long_name_%d:UMI_ACACACAC" % i, "ACACACAC", rd(1800, L, 0)) for i in range(3)]
    prs.append(SynthPair(b"e9s:UMI_ACACACAC", "ACACACAC", rd(1800, L, 0)))
    bb.add(SynthCluster(0, prs, cross_contig=True))
    # 10. threshold 0 vs 1 on the same UMI family (Q18)
    for thr in (0, 1):
        prs = []
        for i in range(4):
            u = "ACGTACGT" if i < 3 else "ACGTACGA"
            prs.append(SynthPair(b"e10_%d_%d:UMI_%s" % (thr, i, u.encode()), u, rd(2000, L, 150), rd(2050, L, -150)))
        bb.add(SynthCluster(0, prs, umi_thr=thr))
    # 11. ties between UMIs: lexicographically first of the max count wins; non-transitive absorption
    prs = []
    for i, u in enumerate(["CCCCCCCC", "CCCCCCCA", "CCCCCCAA", "AAAAAAAA", "CCCCCCCC", "CCCCCCAA"]):
        prs.append(SynthPair(b"e11_%d:UMI_%s" % (i, u.encode()), u, rd(2300, L, 150), rd(2350, L, -150)))
    bb.add(SynthCluster(0, prs))
    # 12. right reads with soft clips / different lengths: right-aligned mode and lenDiff (Q10)
    prs = []
    for i in range(5):
        ll = L - (i % 3)
        l = rd(2600, L, 200)
        r = SynthRead(2700 + (L - ll), f"{ll}M", bytes(ref[2700 + (L - ll):2700 + L]), _q(ll), -200)
        prs.append(SynthPair(b"e12_%d:UMI_GTGTGTGT" % i, "GTGTGTGT", l, r))
    r = SynthRead(2700, f"{L - 5}M5S", bytes(ref[2700:2700 + L - 5]) + b"AAAAA", _q(L), -200)
    prs.append(SynthPair(b"e12_c:UMI_GTGTGTGT", "GTGTGTGT", rd(2600, L, 200), r))
    bb.add(SynthCluster(0, prs))
    # 13. template with insertion + soft clip: getRefOffset == -1 columns; deletion
    prs = []
    for i in range(4):
        seq = bytes(ref[3000:3040]) + b"GG" + bytes(ref[3040:3090]) + b"TTTTTTTT"
        q = _q(100); q = q.copy(); q[40:42] = 11 if i == 0 else 37
        l = SynthRead(3000, "40M2I50M8S", seq if i else mut(seq, [40, 95, 20]), q, 190, 2)
        seq = bytes(ref[3090:3130]) + bytes(ref[3133:3193])
        r = SynthRead(3090, "40M3D60M", seq, _q(100, 25 if i == 1 else 37), -190, 3)
        prs.append(SynthPair(b"e13_%d:UMI_AGAGAGAG" % i, "AGAGAGAG", l, r))
    bb.add(SynthCluster(0, prs))
    # 14. no majority (<40% containment) -> NULL template on one side (group.cpp:264)
    prs = []
    for i, cg in enumerate(["100M", "50M50S", "30M70S", "20S80M", "10S90M", "60M40S"]):
        cgops = synth.parse_cigar(cg)
        l = SynthRead(3400, cg, bytes(ref[3400:3500]), _q(100), 150)
        prs.append(SynthPair(b"e14_%d:UMI_CACACACA" % i, "CACACACA", l, rd(3450, L, -150)))
    bb.add(SynthCluster(0, prs))
    # 15. all-N / qual-0 columns, code 0 ('=') bases, quals >= 128 (signed char refBaseQual quirk)
    prs = []
    for i in range(3):
        l = rd(3700, L, 150); r = rd(3750, L, -150)
        l.seq = mut(l.seq, [1, 2], "N"); l.qual = l.qual.copy(); l.qual[[1, 2]] = 0
        if i == 0:
            l.seq = mut(l.seq, [4], "="); l.qual[6] = 200; l.seq = mut(l.seq, [6])
        if i == 1:
            l.qual[6] = 10; l.qual[8] = 255
        prs.append(SynthPair(b"e15_%d:UMI_TGTGTGTG" % i, "TGTGTGTG", l, r))
    bb.add(SynthCluster(0, prs))
    # 16. qname length rule: left template has the longer padded name -> right's name copied over left (group.cpp:115-122)
    prs = [SynthPair(b"e16_a_longer_name:UMI_GAGAGAGA", "GAGAGAGA", rd(4000, L - 1, 150), rd(4050, L, -150)),
           SynthPair(b"e16_b:UMI_GAGAGAGA", "GAGAGAGA", rd(4000, L, 150), rd(4050, L - 1, -150)),
           SynthPair(b"e16_c:UMI_GAGAGAGA", "GAGAGAGA", rd(4000, L, 150), rd(4050, L, -150))]
    bb.add(SynthCluster(0, prs))
    # 17. read reaching the contig end: pos+len >= contigLen disables the reference (reference.cpp:60)
    end = len(ref)
    prs = []
    for i in range(2):
        l = rd(end - 160, L, 160); r = rd(end - L, L, -160)
        if i == 0:
            r.seq = mut(r.seq, [L - 3]); r.qual = r.qual.copy(); r.qual[L - 3] = 2
        prs.append(SynthPair(b"e17_%d:UMI_ATATATAT" % i, "ATATATAT", l, r))
    bb.add(SynthCluster(0, prs))
    # 18. isize == 0 disables the reference (group.cpp:363); cigar-less reads
    prs = []
    for i in range(2):
        l = rd(4300, L, 0); r = rd(4350, L, 0)
        if i == 0:
            l.seq = mut(l.seq, [9]); l.qual = l.qual.copy(); l.qual[9] = 2
        prs.append(SynthPair(b"e18_%d:UMI_CGCGCGCG" % i, "CGCGCGCG", l, r))
    bb.add(SynthCluster(0, prs))
    # 19. duplex pair where one strand lacks the right read; A_ / _A style UMIs never pair
    bb.add(SynthCluster(0, [SynthPair(b"e19_a:UMI_%s_%s" % (A.encode(), B.encode()), A + "_" + B, rd(4500, L, 170), rd(4570, L, -170)),
                            SynthPair(b"e19_b:UMI_%s_%s" % (B.encode(), A.encode()), B + "_" + A, rd(4500, L - 2, 170))]))
    bb.add(SynthCluster(0, [SynthPair(b"e19_c:UMI_ACGT_", "ACGT_", rd(4600, L, 170), rd(4670, L, -170)),
                            SynthPair(b"e19_d:UMI__ACGT", "_ACGT", rd(4600, L, 170), rd(4670, L, -170))]))
    # 20. second contig
    ref2 = contigs[1]
    prs = []
    for i in range(3):
        sl = bytes(ref2[100:200]); sr = bytes(ref2[160:260])
        if i == 0:
            sl = mut(sl, [50])
        prs.append(SynthPair(b"e20_%d:UMI_TCTCTCTC" % i, "TCTCTCTC", SynthRead(100, "100M", sl, _q(100, 11 if i == 0 else 37), 160, 0),
                             SynthRead(160, "100M", sr, _q(100), -160, 0)))
    bb.add(SynthCluster(1, prs))
    return bb.build(), genome, contigs


def deep_batch(seed: int = 11, big: int = 1100):
    """One family above skipLowComplexityClusterThreshold (1000): low-complexity skip + early break (group.cpp:142-175,231)."""
    return synth.make_ragged_batch(seed, n_clusters=6, depth=5.0, umi="single", big_cluster=big, indel_frac=0.3, clip_frac=0.4)


def low_complexity_batch(seed: int = 13, n: int = 1050):
    """>1000 pairs, >10% distinct CIGARs, homopolymer-ish first read -> both sides skipped."""
    rng = np.random.Generator(np.random.PCG64(seed))
    contigs, genome = synth.random_genome(rng, [4000])
    contigs[0][500:900] = ord("A")
    genome = synth.Genome.from_bases(contigs)
    bb = BatchBuilder(1, "UMI")
    prs = []
    for i in range(n):
        k = int(rng.integers(1, 60))
        l = SynthRead(500, f"{100 - k}M{k}S", bytes(contigs[0][500:600]), _q(100), 180)
        r = SynthRead(580, f"{k}S{100 - k}M", bytes(contigs[0][580:680]), _q(100), -180)
        prs.append(SynthPair(b"lc_%05d:UMI_ACGTACGT" % i, "ACGTACGT", l, r))
    bb.add(SynthCluster(0, prs))
    return bb.build(), genome, contigs


OPTION_SETS = {
    "default": Options.default(),
    "s2_x": Options.default(cluster_size_req=2, duplex_only=1),
    "noduplex_D0": Options.default(disable_duplex=1, duplex_mismatch_threshold=0),
    "strict": Options.default(cluster_size_req=3, base_score_req=10, score_percent_req=1.0, high_quality=40, moderate_quality=35,
                              low_quality=30, duplex_mismatch_threshold=0),
    "loose": Options.default(base_score_req=1, score_percent_req=0.5, high_quality=20, moderate_quality=15, low_quality=8,
                             duplex_mismatch_threshold=10),
}
