"""Synthetic workloads in the packed batch format (SURVEY §8(d) generator, BASELINE.json configs).

Three producers:
  * `BatchBuilder`   — read-by-read construction (tests, edge cases, ragged/clip/indel stress);
  * `make_fixed_batch` — vectorised numpy generator for the fixed-length 2x150 configs
    (cfg1..cfg4 shapes) that scales to millions of pairs in seconds (bench.py);
  * `make_ragged_fixed` — vectorised generator of the cfg5 shape (mixed lengths, soft clips, indels);
  `make_batch` picks between the last two by the config.
All emit pairs of a cluster in the iteration order of the reference's map<qname+NUL padding>.
No consensus arithmetic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .abi import (CLUSTER_CROSS_CONTIG, CLUSTER_UMI_THR_SHIFT, READ_DESC, Batch, Genome, align4, encode_umi,
                  padded_l_qname)

BASES = np.frombuffer(b"ACGT", np.uint8)
BAM_CODE = np.zeros(256, np.uint8)  # ASCII -> BAM 4-bit (bamutil.cpp:167-183); others -> N
BAM_CODE[:] = 15
for _ch, _v in ((b"A", 1), (b"C", 2), (b"G", 4), (b"T", 8), (b"N", 15), (b"=", 0)):
    BAM_CODE[_ch[0]] = _v
CIGAR_OPS = "MIDNSHP=XB"
QUERY_CONSUMING = set("MIS=X")
REF_CONSUMING = set("MDN=X")


def pack_bases(codes: np.ndarray) -> np.ndarray:
    """BAM 4-bit codes (1-D) -> packed bytes, even index in the high nibble."""
    if len(codes) % 2:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def parse_cigar(s: str) -> np.ndarray:
    ops, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            ops.append((int(num) << 4) | CIGAR_OPS.index(ch))
            num = ""
    return np.asarray(ops, np.uint32)


def cigar_query_len(c: np.ndarray) -> int:
    return int(sum(int(v) >> 4 for v in c if CIGAR_OPS[int(v) & 15] in QUERY_CONSUMING))


def random_genome(rng: np.random.Generator, contig_lens: Sequence[int]) -> Tuple[List[np.ndarray], Genome]:
    contigs = [BASES[rng.integers(0, 4, n, dtype=np.uint8)] for n in contig_lens]
    return contigs, Genome.from_bases(contigs)


@dataclass
class SynthRead:
    pos: int
    cigar: str
    seq: bytes            # ASCII bases, len == query length of cigar
    qual: np.ndarray      # uint8 phred
    isize: int = 0
    nm: int = 0


@dataclass
class SynthPair:
    qname: bytes
    umi: str
    left: SynthRead
    right: Optional[SynthRead] = None


@dataclass
class SynthCluster:
    ref: int                      # genome contig index or -1
    pairs: List[SynthPair] = field(default_factory=list)
    umi_thr: int = 1
    cross_contig: bool = False


class BatchBuilder:
    def __init__(self, umi_words: int = 1, umi_prefix: str = ""):
        self.clusters: List[SynthCluster] = []
        self.umi_words = umi_words
        self.umi_prefix = umi_prefix

    def add(self, c: SynthCluster) -> SynthCluster:
        self.clusters.append(c)
        return c

    def build(self) -> Batch:
        n_pairs = sum(len(c.pairs) for c in self.clusters)
        pair_off = np.zeros(len(self.clusters) + 1, np.int32)
        cref = np.zeros(len(self.clusters), np.int32)
        cflags = np.zeros(len(self.clusters), np.uint8)
        umi = np.zeros((n_pairs, self.umi_words), np.uint64)
        reads = np.zeros(2 * n_pairs, READ_DESC)
        reads["l_qseq"] = -1
        nm = np.zeros(2 * n_pairs, np.uint8)
        cigars: List[np.ndarray] = []
        n_cig = 0
        chunks: List[np.ndarray] = []
        cursor = 0
        qnames: List[bytes] = []
        p = 0
        for ci, c in enumerate(self.clusters):
            # iteration order of map<string(qname, l_qname incl. NUL padding)> (bamutil.cpp:19-21)
            def key(pr: SynthPair) -> bytes:
                return pr.qname + b"\0" * (padded_l_qname(len(pr.qname)) - len(pr.qname))
            pairs = sorted(c.pairs, key=key)
            assert len({key(x) for x in pairs}) == len(pairs), "duplicate qname in a cluster"
            pair_off[ci] = p
            cref[ci] = c.ref
            cflags[ci] = (CLUSTER_CROSS_CONTIG if c.cross_contig else 0) | (c.umi_thr << CLUSTER_UMI_THR_SHIFT)
            pad = (-cursor) % 16
            if pad:
                chunks.append(np.zeros(pad, np.uint8))
                cursor += pad
            for pr in pairs:
                qnames.append(pr.qname)
                umi[p] = encode_umi(pr.umi, self.umi_words)
                for s, r in enumerate((pr.left, pr.right)):
                    if r is None:
                        continue
                    cg = parse_cigar(r.cigar)
                    l = len(r.seq)
                    assert len(cg) == 0 or cigar_query_len(cg) == l, (r.cigar, l)
                    assert len(r.qual) == l
                    d = reads[2 * p + s]
                    d["data_off"], d["l_qseq"], d["pos"], d["isize"] = cursor, l, r.pos, r.isize
                    d["cigar_off"], d["n_cigar"], d["l_qname"] = n_cig, len(cg), padded_l_qname(len(pr.qname))
                    nm[2 * p + s] = min(r.nm, 255)
                    cigars.append(cg)
                    n_cig += len(cg)
                    rec = np.zeros(align4(l) + align4((l + 1) // 2), np.uint8)
                    rec[:l] = r.qual
                    rec[align4(l):align4(l) + (l + 1) // 2] = pack_bases(BAM_CODE[np.frombuffer(r.seq, np.uint8)])
                    chunks.append(rec)
                    cursor += len(rec)
                p += 1
        pair_off[-1] = p
        pad = (-cursor) % 16
        if pad:
            chunks.append(np.zeros(pad, np.uint8))
        payload = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
        if len(payload) == 0:
            payload = np.zeros(16, np.uint8)
        cigar = np.concatenate(cigars).astype(np.uint32) if cigars else np.zeros(0, np.uint32)
        if len(cigar) == 0:
            cigar = np.zeros(1, np.uint32)
        b = Batch(pair_off, cref, cflags, umi, reads, cigar, np.ascontiguousarray(payload), qnames, nm, self.umi_prefix)
        b.validate()
        return b


# ----------------------------------------------------------------------------- ragged generator


def _qual_draw(rng, n, low_bias=False):
    p = [0.3, 0.3, 0.3, 0.1] if low_bias else [0.90, 0.06, 0.03, 0.01]
    return rng.choice(np.asarray([37, 25, 11, 2], np.uint8), size=n, p=p)


def make_ragged_batch(seed: int, n_clusters: int = 200, depth: float = 6.0, read_len=(100, 250), umi: str = "single",
                      err: float = 0.01, umi_err: float = 0.02, clip_frac: float = 0.2, indel_frac: float = 0.05,
                      missing_mate_frac: float = 0.03, cross_frac: float = 0.03, multi_frac: float = 0.2,
                      n_frac: float = 0.002, contig_len: int = 200_000, n_contigs: int = 2, absent_ref_frac: float = 0.03,
                      tail_thr_frac: float = 0.2, same_right_pos_frac: float = 0.1, hard_clip_frac: float = 0.05,
                      big_cluster: int = 0) -> Tuple[Batch, Genome, List[np.ndarray]]:
    """cfg5-style stress: mixed lengths, soft/hard clips, indels, missing mates, cross-contig clusters,
    multi-family clusters, N bases, clusters on a contig the FASTA lacks, both UMI thresholds."""
    rng = np.random.Generator(np.random.PCG64(seed))
    contigs, genome = random_genome(rng, [contig_len] * n_contigs)
    words = 2 if umi == "duplex" else 1
    bb = BatchBuilder(words, "UMI" if umi != "none" else "")
    lo, hi = read_len if isinstance(read_len, tuple) else (read_len, read_len)
    mol_id = 0

    def rand_umi(k=8):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, k))

    def mutate_umi(u):
        return "".join(("ACGT"[rng.integers(0, 4)] if (ch != "_" and rng.random() < umi_err) else ch) for ch in u)

    def make_read(contig, pos, l, insert_isize, allow_shape=True):
        """A read of query length l whose first aligned base is at reference pos."""
        ops = []
        soft_l = soft_r = 0
        if allow_shape and rng.random() < clip_frac:
            if rng.random() < 0.5:
                soft_l = int(rng.integers(1, min(30, l // 3)))
            else:
                soft_r = int(rng.integers(1, min(30, l // 3)))
        core = l - soft_l - soft_r
        ref = contigs[contig]
        seq = bytearray()
        if allow_shape and rng.random() < hard_clip_frac:
            ops.append(f"{int(rng.integers(1, 20))}H")
        if soft_l:
            ops.append(f"{soft_l}S")
            seq += bytes(BASES[rng.integers(0, 4, soft_l)])
        rpos = pos
        nm = 0
        if allow_shape and rng.random() < indel_frac and core > 40:
            a = int(rng.integers(10, core - 20))
            k = int(rng.integers(1, 11))
            if rng.random() < 0.5:  # insertion
                k = min(k, core - a - 5)
                ops += [f"{a}M", f"{k}I", f"{core - a - k}M"]
                seq += bytes(ref[rpos:rpos + a]) + bytes(BASES[rng.integers(0, 4, k)]) + bytes(ref[rpos + a:rpos + core - k])
            else:
                ops += [f"{a}M", f"{k}D", f"{core - a}M"]
                seq += bytes(ref[rpos:rpos + a]) + bytes(ref[rpos + a + k:rpos + k + core])
            nm += k
        else:
            ops.append(f"{core}M")
            seq += bytes(ref[rpos:rpos + core])
        if soft_r:
            ops.append(f"{soft_r}S")
            seq += bytes(BASES[rng.integers(0, 4, soft_r)])
        assert len(seq) == l, (len(seq), l, ops)
        return "".join(ops), bytes(seq), nm

    def noisy_copy(seq: bytes, nm: int):
        a = np.frombuffer(seq, np.uint8).copy()
        q = _qual_draw(rng, len(a))
        e = rng.random(len(a)) < err
        k = int(e.sum())
        if k:
            a[e] = BASES[(np.searchsorted(BASES, a[e]) + rng.integers(1, 4, k)) % 4]
            q[e] = _qual_draw(rng, k, low_bias=True)
        nmask = rng.random(len(a)) < n_frac
        a[nmask] = ord("N")
        q[nmask] = 2
        return bytes(a), q, nm + k

    for _ in range(n_clusters):
        contig = int(rng.integers(0, n_contigs))
        cross = rng.random() < cross_frac
        cl = SynthCluster(ref=(-1 if rng.random() < absent_ref_frac else contig),
                          umi_thr=(0 if rng.random() < tail_thr_frac else 1), cross_contig=cross)
        n_mol = 1 + (int(rng.integers(1, 3)) if rng.random() < multi_frac else 0)
        lmax = hi
        insert = int(np.clip(rng.normal(1.45 * (lo + hi) / 2, 40), hi + 5, 2 * hi + 200))
        start = int(rng.integers(50, contig_len - insert - 400))
        same_right = rng.random() < same_right_pos_frac
        for _m in range(n_mol):
            mol_id += 1
            fam = 1 + int(rng.poisson(max(depth - 1, 0.0)))
            if big_cluster and _ == 0 and _m == 0:
                fam = big_cluster
            ua, ub = rand_umi(), rand_umi()
            # fragment-level template shapes shared by most family members (so isPartOf finds a majority)
            ll = int(rng.integers(lo, hi + 1))
            rl = int(rng.integers(lo, hi + 1))
            lc, lseq, lnm = make_read(contig, start, ll, insert)
            rpos = start + insert - rl if not same_right else start + insert - hi
            rc, rseq, rnm = make_read(contig, rpos, rl, -insert)
            for d in range(fam):
                strand_top = True if umi != "duplex" else bool(rng.random() < 0.5)
                if umi == "none":
                    u, tail = "", b""
                elif umi == "single":
                    u = mutate_umi(ua)
                    tail = b":UMI_" + u.encode()
                else:
                    u = mutate_umi(ua + "_" + ub if strand_top else ub + "_" + ua)
                    tail = b":UMI_" + u.encode()
                # a minority of family members get their own length/shape (ragged columns)
                if rng.random() < 0.3:
                    l2 = int(rng.integers(lo, hi + 1))
                    c2, s2, n2 = make_read(contig, start, l2, insert)
                else:
                    c2, s2, n2 = lc, lseq, lnm
                if rng.random() < 0.3:
                    r2 = int(rng.integers(lo, hi + 1))
                    rp2 = start + insert - r2 if not same_right else rpos
                    rc2, rs2, rn2 = make_read(contig, rp2, r2, -insert)
                else:
                    rp2, rc2, rs2, rn2 = rpos, rc, rseq, rnm
                sq, ql, nmv = noisy_copy(s2, n2)
                left = SynthRead(start, c2, sq, ql, insert, nmv)
                right = None
                if not cross and rng.random() >= missing_mate_frac:
                    sq, ql, nmv = noisy_copy(rs2, rn2)
                    right = SynthRead(rp2, rc2, sq, ql, -insert, nmv)
                # names of varying length exercise the padded-l_qname rules (group.cpp:90-96,115-122)
                stem = b"SIM:%d:%d" % (mol_id, d) if rng.random() < 0.5 else b"SIM:%06d:%03d" % (mol_id, d)
                cl.pairs.append(SynthPair(stem + tail, u, left, right))
        bb.add(cl)
    return bb.build(), genome, contigs


# ----------------------------------------------------------------------------- vectorised generator


@dataclass
class FixedConfig:
    """The shapes of BASELINE.json configs 0-4 (cfg1..cfg5 in SURVEY §8(d)); cfg5 is the ragged one (`make_batch`)."""
    name: str
    n_pairs: int
    read_len: int = 150
    depth: float = 8.0            # mean pairs per molecule
    umi: str = "single"           # none | single | duplex
    err: float = 0.001
    umi_err: float = 0.01
    shared_frac: float = 0.10     # molecules sharing (start,end) with the previous one
    n_contigs: int = 1
    contig_len: int = 50_000_000
    insert_mu: float = 220.0
    insert_sigma: float = 40.0
    umi_thr: int = 1
    supporting_reads: int = 1
    # cfg5 (ragged-column stress): reads of read_len_min..read_len bases, a soft clip of 1-30 bases at one end of
    # clip_frac of the reads, one 1-10 base insertion or deletion in indel_frac of them
    read_len_min: int = 0         # 0 = fixed length
    clip_frac: float = 0.0
    indel_frac: float = 0.0


CONFIGS: Dict[str, FixedConfig] = {
    "cfg1": FixedConfig("cfg1", 10_000, depth=4.0, umi="none", err=0.001, shared_frac=0.0, n_contigs=1, contig_len=1_000_000),
    "cfg2": FixedConfig("cfg2", 1_000_000, depth=8.0, umi="single", err=0.001, umi_err=0.01, shared_frac=0.10,
                        n_contigs=1, contig_len=50_000_000),
    "cfg3": FixedConfig("cfg3", 10_000_000, depth=20.0, umi="duplex", err=0.001, n_contigs=16, contig_len=10_000_000),
    "cfg4": FixedConfig("cfg4", 50_000_000, depth=100.0, umi="duplex", err=0.01, n_contigs=16, contig_len=10_000_000,
                        insert_mu=167.0, insert_sigma=20.0, supporting_reads=2),
    "cfg5": FixedConfig("cfg5", 5_000_000, read_len=250, read_len_min=100, depth=6.0, umi="single", err=0.001, umi_err=0.01,
                        shared_frac=0.10, n_contigs=4, contig_len=25_000_000, insert_mu=330.0, insert_sigma=60.0,
                        clip_frac=0.20, indel_frac=0.05),
}

_QUAL_LUT = np.empty(256, np.uint8)  # {37:.90, 25:.06, 11:.03, 2:.01} on a 1/256 grid
_QUAL_LUT[:230] = 37
_QUAL_LUT[230:245] = 25
_QUAL_LUT[245:253] = 11
_QUAL_LUT[253:] = 2
_ERRQ_LUT = np.empty(256, np.uint8)  # error bases biased to the low bins
_ERRQ_LUT[:77] = 37
_ERRQ_LUT[77:154] = 25
_ERRQ_LUT[154:231] = 11
_ERRQ_LUT[231:] = 2


def make_fixed_batch(cfg: FixedConfig, seed: int, n_pairs: Optional[int] = None, with_qnames: bool = True,
                     genome_cache: Optional[Tuple[List[np.ndarray], Genome]] = None):
    """Returns (Batch, Genome, contigs).  Deterministic in (cfg, seed, n_pairs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = cfg.read_len
    n_target = cfg.n_pairs if n_pairs is None else n_pairs
    if genome_cache is None:
        contigs, genome = random_genome(rng, [cfg.contig_len] * cfg.n_contigs)
    else:
        contigs, genome = genome_cache
    # ---- molecules
    n_mol = int(n_target / cfg.depth * 1.25) + 64
    fam = 1 + rng.poisson(max(cfg.depth - 1.0, 0.0), n_mol).astype(np.int64)
    while fam.sum() < n_target:
        fam = np.concatenate([fam, 1 + rng.poisson(max(cfg.depth - 1.0, 0.0), n_mol).astype(np.int64)])
    csum = np.cumsum(fam)
    n_mol = int(np.searchsorted(csum, n_target, side="left")) + 1
    fam = fam[:n_mol]
    fam[-1] -= int(csum[n_mol - 1] - n_target)
    if fam[-1] <= 0:
        fam, n_mol = fam[:-1], n_mol - 1
    N = int(fam.sum())
    insert = np.clip(np.rint(rng.normal(cfg.insert_mu, cfg.insert_sigma, n_mol)), L, 2 * L + 100).astype(np.int64)
    contig = rng.integers(0, cfg.n_contigs, n_mol)
    start = rng.integers(0, cfg.contig_len - insert.max() - 1, n_mol)
    shared = np.flatnonzero(rng.random(n_mol) < cfg.shared_frac)
    shared = shared[shared > 0]
    for a in (insert, contig, start):  # copy coordinates of the previous molecule (chains allowed)
        a[shared] = a[shared - 1]
    order = np.lexsort((np.arange(n_mol), insert, start, contig))
    fam, insert, contig, start = fam[order], insert[order], contig[order], start[order]
    new_cluster = np.ones(n_mol, bool)
    new_cluster[1:] = (contig[1:] != contig[:-1]) | (start[1:] != start[:-1]) | (insert[1:] != insert[:-1])
    cluster_of_mol = np.cumsum(new_cluster) - 1
    n_clusters = int(cluster_of_mol[-1]) + 1
    # ---- pairs (molecule-major, dup index minor == qname order)
    mol_of_pair = np.repeat(np.arange(n_mol), fam)
    pair_first = np.concatenate([[0], np.cumsum(fam)[:-1]])
    dup = np.arange(N) - pair_first[mol_of_pair]
    cluster_of_pair = cluster_of_mol[mol_of_pair]
    cluster_pair_off = np.zeros(n_clusters + 1, np.int32)
    np.add.at(cluster_pair_off, cluster_of_pair + 1, 1)
    cluster_pair_off = np.cumsum(cluster_pair_off).astype(np.int32)
    first_mol_of_cluster = np.flatnonzero(new_cluster)
    cluster_ref = contig[first_mol_of_cluster].astype(np.int32)
    cluster_flags = np.full(n_clusters, cfg.umi_thr << CLUSTER_UMI_THR_SHIFT, np.uint8)
    # ---- UMIs
    if cfg.umi == "none":
        words, umi = 1, np.zeros((N, 1), np.uint64)
        umi_chars = None
    else:
        k = 8
        nfield = k if cfg.umi == "single" else 2 * k + 1
        words = (nfield + 15) // 16
        ua = rng.integers(0, 4, (n_mol, k), dtype=np.uint8)
        fields = np.zeros((N, nfield), np.uint8)
        if cfg.umi == "single":
            fields[:] = ua[mol_of_pair] + 1
        else:
            ub = rng.integers(0, 4, (n_mol, k), dtype=np.uint8)
            top = rng.random(N) < 0.5
            a, b_ = ua[mol_of_pair] + 1, ub[mol_of_pair] + 1
            fields[:, :k] = np.where(top[:, None], a, b_)
            fields[:, k] = 5
            fields[:, k + 1:] = np.where(top[:, None], b_, a)
        nerr = rng.binomial(N * nfield, cfg.umi_err)
        if nerr:
            ei = rng.integers(0, N, nerr)
            ej = rng.integers(0, nfield, nerr)
            ok = fields[ei, ej] != 5
            fields[ei[ok], ej[ok]] = rng.integers(1, 5, int(ok.sum()), dtype=np.uint8)
        umi = np.zeros((N, words), np.uint64)
        for j in range(nfield):
            umi[:, j >> 4] |= fields[:, j].astype(np.uint64) << np.uint64(60 - 4 * (j & 15))
        umi_chars = np.frombuffer(b"?ACGT_", np.uint8)[fields]
    # ---- reads: slot 2p = left, 2p+1 = right
    goff = np.concatenate([[0], np.cumsum([len(c) for c in contigs])[:-1]]).astype(np.int64)
    gall = np.concatenate(contigs) if len(contigs) > 1 else contigs[0]
    lpos = start[mol_of_pair]
    rpos = lpos + insert[mol_of_pair] - L
    pos = np.empty(2 * N, np.int64)
    pos[0::2], pos[1::2] = lpos, rpos
    gbase = goff[contig[mol_of_pair]]
    gstart = np.repeat(gbase, 2) + pos
    ar = np.arange(L, dtype=np.int64)
    seq = gall[gstart[:, None] + ar[None, :]]          # (2N, L) ASCII
    qual = _QUAL_LUT[rng.integers(0, 256, (2 * N, L), dtype=np.uint8)]
    nerr = rng.binomial(2 * N * L, cfg.err)
    nm = np.zeros(2 * N, np.int64)
    if nerr:
        ei = rng.integers(0, 2 * N, nerr)
        ej = rng.integers(0, L, nerr)
        old = np.searchsorted(BASES, seq[ei, ej])
        seq[ei, ej] = BASES[(old + rng.integers(1, 4, nerr)) % 4]
        qual[ei, ej] = _ERRQ_LUT[rng.integers(0, 256, nerr, dtype=np.uint8)]
        np.add.at(nm, ei, 1)
    codes = BAM_CODE[seq]
    if L % 2:
        codes = np.concatenate([codes, np.zeros((2 * N, 1), np.uint8)], axis=1)
    packed = (codes[:, 0::2] << 4) | codes[:, 1::2]
    QB, SB = align4(L), align4((L + 1) // 2)
    rec = QB + SB
    recs = np.zeros((2 * N, rec), np.uint8)
    recs[:, :L] = qual
    recs[:, QB:QB + (L + 1) // 2] = packed
    del seq, qual, codes, packed
    reads_per_cluster = 2 * np.diff(cluster_pair_off).astype(np.int64)
    cbytes = (reads_per_cluster * rec + 15) & ~15
    cstart = np.concatenate([[0], np.cumsum(cbytes)[:-1]])
    total = int(cbytes.sum())
    idx_in_cluster = np.arange(2 * N) - 2 * np.repeat(cluster_pair_off[:-1].astype(np.int64), reads_per_cluster)
    data_off = np.repeat(cstart, reads_per_cluster) + idx_in_cluster * rec
    payload = np.zeros(total, np.uint8)
    # records are 4-byte multiples at 4-byte aligned offsets: scatter as u32 rows
    p32 = payload.view(np.uint32)
    p32[(data_off // 4)[:, None] + np.arange(rec // 4)[None, :]] = recs.view(np.uint32)
    del recs
    reads = np.zeros(2 * N, READ_DESC)
    reads["data_off"] = data_off
    reads["l_qseq"] = L
    reads["pos"] = pos
    isz = np.repeat(insert[mol_of_pair], 2)
    isz[1::2] *= -1
    reads["isize"] = isz
    reads["cigar_off"] = 0
    reads["n_cigar"] = 1
    cigar = np.asarray([(L << 4) | 0], np.uint32)
    # ---- qnames: SIM:<mol 9 digits>:<dup 5 digits>[:UMI_<umi>]  (fixed width => numeric order == string order)
    qn = None
    width = 4 + 9 + 1 + 5 + (0 if umi_chars is None else 5 + umi_chars.shape[1])
    if with_qnames:
        mat = np.empty((N, width), np.uint8)
        mat[:, 0:4] = np.frombuffer(b"SIM:", np.uint8)
        m = mol_of_pair.copy()
        for j in range(9):
            mat[:, 12 - j] = 48 + (m % 10)
            m //= 10
        mat[:, 13] = ord(":")
        d = dup.copy()
        for j in range(5):
            mat[:, 18 - j] = 48 + (d % 10)
            d //= 10
        if umi_chars is not None:
            mat[:, 19:24] = np.frombuffer(b":UMI_", np.uint8)
            mat[:, 24:] = umi_chars
        qn = np.ascontiguousarray(mat).view(f"S{width}").ravel()
    reads["l_qname"] = padded_l_qname(width)
    b = Batch(cluster_pair_off, cluster_ref, cluster_flags, umi, reads, cigar, payload, qn,
              np.minimum(nm, 255).astype(np.uint8), "UMI" if cfg.umi != "none" else "")
    b.validate()
    return b, genome, contigs


# ----------------------------------------------------------------------------- vectorised ragged generator (cfg5)


def make_ragged_fixed(cfg: FixedConfig, seed: int, n_pairs: Optional[int] = None, with_qnames: bool = True,
                      genome_cache: Optional[Tuple[List[np.ndarray], Genome]] = None, own_len_frac: float = 0.3):
    """cfg5: read length uniform read_len_min..read_len per molecule and side, `own_len_frac` of the reads trimmed
    further at their 3' end (left reads keep their start, right reads their end: the right-aligned column mode of group.cpp:339-349),
    clip_frac of the reads with a 1-30 base soft clip at one end, indel_frac with one 1-10 base insertion or deletion.
    Returns (Batch, Genome, contigs); deterministic in (cfg, seed, n_pairs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Lmin, Lmax = cfg.read_len_min, cfg.read_len
    n_target = cfg.n_pairs if n_pairs is None else n_pairs
    if genome_cache is None:
        contigs, genome = random_genome(rng, [cfg.contig_len] * cfg.n_contigs)
    else:
        contigs, genome = genome_cache
    # ---- molecules (as make_fixed_batch)
    n_mol = int(n_target / cfg.depth * 1.25) + 64
    fam = 1 + rng.poisson(max(cfg.depth - 1.0, 0.0), n_mol).astype(np.int64)
    while fam.sum() < n_target:
        fam = np.concatenate([fam, 1 + rng.poisson(max(cfg.depth - 1.0, 0.0), n_mol).astype(np.int64)])
    csum = np.cumsum(fam)
    n_mol = int(np.searchsorted(csum, n_target, side="left")) + 1
    fam = fam[:n_mol]
    fam[-1] -= int(csum[n_mol - 1] - n_target)
    if fam[-1] <= 0:
        fam, n_mol = fam[:-1], n_mol - 1
    N = int(fam.sum())
    mol_len = rng.integers(Lmin, Lmax + 1, (n_mol, 2))
    insert = np.rint(rng.normal(cfg.insert_mu, cfg.insert_sigma, n_mol)).astype(np.int64)
    insert = np.clip(np.maximum(insert, mol_len.max(axis=1) + 12), Lmin, 2 * Lmax + 100)
    contig = rng.integers(0, cfg.n_contigs, n_mol)
    start = rng.integers(40, cfg.contig_len - insert.max() - 80, n_mol)
    shared = np.flatnonzero(rng.random(n_mol) < cfg.shared_frac)
    shared = shared[shared > 0]
    for a in (insert, contig, start):
        a[shared] = a[shared - 1]
    order = np.lexsort((np.arange(n_mol), insert, start, contig))
    fam, insert, contig, start, mol_len = fam[order], insert[order], contig[order], start[order], mol_len[order]
    mol_len = np.minimum(mol_len, (insert - 12)[:, None])  # (a molecule that took another's coordinates keeps its reads inside them)
    new_cluster = np.ones(n_mol, bool)
    new_cluster[1:] = (contig[1:] != contig[:-1]) | (start[1:] != start[:-1]) | (insert[1:] != insert[:-1])
    cluster_of_mol = np.cumsum(new_cluster) - 1
    n_clusters = int(cluster_of_mol[-1]) + 1
    mol_of_pair = np.repeat(np.arange(n_mol), fam)
    pair_first = np.concatenate([[0], np.cumsum(fam)[:-1]])
    dup = np.arange(N) - pair_first[mol_of_pair]
    cluster_of_pair = cluster_of_mol[mol_of_pair]
    cluster_pair_off = np.zeros(n_clusters + 1, np.int32)
    np.add.at(cluster_pair_off, cluster_of_pair + 1, 1)
    cluster_pair_off = np.cumsum(cluster_pair_off).astype(np.int32)
    cluster_ref = contig[np.flatnonzero(new_cluster)].astype(np.int32)
    cluster_flags = np.full(n_clusters, cfg.umi_thr << CLUSTER_UMI_THR_SHIFT, np.uint8)
    # ---- UMIs (8 characters, one word)
    k = 8
    fields = (rng.integers(0, 4, (n_mol, k), dtype=np.uint8) + 1)[mol_of_pair]
    nerr = rng.binomial(N * k, cfg.umi_err)
    if nerr:
        fields[rng.integers(0, N, nerr), rng.integers(0, k, nerr)] = rng.integers(1, 5, nerr, dtype=np.uint8)
    umi = np.zeros((N, 1), np.uint64)
    for j in range(k):
        umi[:, 0] |= fields[:, j].astype(np.uint64) << np.uint64(60 - 4 * j)
    umi_chars = np.frombuffer(b"?ACGT_", np.uint8)[fields]
    # ---- per-read shape: slot 2p = left, 2p+1 = right
    R = 2 * N
    side = np.tile(np.asarray([0, 1]), N)
    l = mol_len[np.repeat(mol_of_pair, 2), side].astype(np.int64)
    own = rng.random(R) < own_len_frac
    l[own] = rng.integers(Lmin, l[own] + 1)
    clip = rng.random(R) < cfg.clip_frac
    clip_n = np.where(clip, rng.integers(1, 31, R), 0)
    clip_left = rng.random(R) < 0.5
    clipL, clipR = np.where(clip_left, clip_n, 0), np.where(clip_left, 0, clip_n)
    core = l - clipL - clipR
    ind = (rng.random(R) < cfg.indel_frac) & (core > 40)
    ind_k = np.where(ind, rng.integers(1, 11, R), 0)
    ind_ins = ind & (rng.random(R) < 0.5)
    ind_del = ind & ~ind_ins
    ind_a = np.where(ind, rng.integers(10, np.maximum(core - 20, 11)), 0)
    refspan = core - np.where(ind_ins, ind_k, 0) + np.where(ind_del, ind_k, 0)
    mstart = np.repeat(start[mol_of_pair], 2)
    mins = np.repeat(insert[mol_of_pair], 2)
    pos = np.where(side == 0, mstart, mstart + mins - refspan)
    goff = np.concatenate([[0], np.cumsum([len(c) for c in contigs])[:-1]]).astype(np.int64)
    gall = np.concatenate(contigs) if len(contigs) > 1 else contigs[0]
    gpos = goff[np.repeat(contig[mol_of_pair], 2)] + pos
    # ---- CIGARs: [clipL S] a M [k I|D] rest M [clipR S]
    ops = np.zeros((R, 5), np.uint32)
    valid = np.zeros((R, 5), bool)
    ops[:, 0], valid[:, 0] = (clipL << 4) | 4, clipL > 0
    first_m = np.where(ind, ind_a, core)
    ops[:, 1], valid[:, 1] = (first_m << 4) | 0, True
    ops[:, 2], valid[:, 2] = (ind_k << 4) | np.where(ind_ins, 1, 2), ind
    ops[:, 3], valid[:, 3] = ((core - ind_a - np.where(ind_ins, ind_k, 0)) << 4) | 0, ind
    ops[:, 4], valid[:, 4] = (clipR << 4) | 4, clipR > 0
    n_cigar = valid.sum(axis=1)
    cigar = ops[valid].astype(np.uint32)
    cigar_off = np.concatenate([[0], np.cumsum(n_cigar)[:-1]])
    # ---- payload layout
    rec = ((l + 3) & ~3) + (((l + 1) // 2 + 3) & ~3)
    reads_per_cluster = 2 * np.diff(cluster_pair_off).astype(np.int64)
    cl_of_read = np.repeat(np.arange(n_clusters), reads_per_cluster)
    rec_csum = np.cumsum(rec)
    cl_first = np.concatenate([[0], np.cumsum(reads_per_cluster)[:-1]])
    cl_bytes = np.add.reduceat(rec, cl_first)
    cl_start = np.concatenate([[0], np.cumsum((cl_bytes + 15) & ~15)[:-1]])
    within = rec_csum - rec - (rec_csum - rec)[cl_first][cl_of_read]
    data_off = cl_start[cl_of_read] + within
    total = int(((cl_bytes + 15) & ~15).sum())
    payload = np.zeros(total, np.uint8)
    nm = np.zeros(R, np.int64)
    W = Lmax + (Lmax & 1)
    ar = np.arange(W, dtype=np.int64)
    CH = 200_000
    for r0 in range(0, R, CH):
        r1 = min(R, r0 + CH)
        n = r1 - r0
        ll, cL, cR, a_, k_, ins_, del_ = l[r0:r1, None], clipL[r0:r1, None], clipR[r0:r1, None], ind_a[r0:r1, None], ind_k[r0:r1, None], \
            ind_ins[r0:r1, None], ind_del[r0:r1, None]
        q = ar[None, :]
        c = q - cL
        inq = q < ll
        rnd = (c < 0) | (q >= ll - cR) | (ins_ & (c >= a_) & (c < a_ + k_))
        shift = np.where(ins_ & (c >= a_ + k_), -k_, 0) + np.where(del_ & (c >= a_), k_, 0)
        gi = np.clip(gpos[r0:r1, None] + c + shift, 0, len(gall) - 1)
        seq = gall[gi]
        rb = BASES[rng.integers(0, 4, (n, W), dtype=np.uint8)]
        seq = np.where(rnd, rb, seq)
        qual = _QUAL_LUT[rng.integers(0, 256, (n, W), dtype=np.uint8)]
        err = (rng.random((n, W)) < cfg.err) & inq & ~rnd
        ne = int(err.sum())
        if ne:
            seq[err] = BASES[(np.searchsorted(BASES, seq[err]) + rng.integers(1, 4, ne)) % 4]
            qual[err] = _ERRQ_LUT[rng.integers(0, 256, ne, dtype=np.uint8)]
        nm[r0:r1] = err.sum(axis=1) + k_[:, 0]
        codes = np.where(inq, BAM_CODE[seq], 0).astype(np.uint8)
        packed = (codes[:, 0::2] << 4) | codes[:, 1::2]
        qual = np.where(inq, qual, 0).astype(np.uint8)
        # scatter: quality bytes, then packed bases at align4(l)
        do = data_off[r0:r1, None]
        qm = inq
        payload[(do + q)[qm]] = qual[qm]
        hb = ar[None, :W // 2]
        sm = hb < (ll + 1) // 2
        payload[(do + ((ll + 3) & ~3) + hb)[sm]] = packed[sm]
    reads = np.zeros(R, READ_DESC)
    reads["data_off"] = data_off
    reads["l_qseq"] = l
    reads["pos"] = pos
    isz = mins.copy()
    isz[1::2] *= -1
    reads["isize"] = isz
    reads["cigar_off"] = cigar_off
    reads["n_cigar"] = n_cigar
    qn = None
    width = 4 + 9 + 1 + 5 + 5 + k
    if with_qnames:
        mat = np.empty((N, width), np.uint8)
        mat[:, 0:4] = np.frombuffer(b"SIM:", np.uint8)
        m = mol_of_pair.copy()
        for j in range(9):
            mat[:, 12 - j] = 48 + (m % 10)
            m //= 10
        mat[:, 13] = ord(":")
        d = dup.copy()
        for j in range(5):
            mat[:, 18 - j] = 48 + (d % 10)
            d //= 10
        mat[:, 19:24] = np.frombuffer(b":UMI_", np.uint8)
        mat[:, 24:] = umi_chars
        qn = np.ascontiguousarray(mat).view(f"S{width}").ravel()
    reads["l_qname"] = padded_l_qname(width)
    b = Batch(cluster_pair_off, cluster_ref, cluster_flags, umi, reads, cigar, payload, qn, np.minimum(nm, 255).astype(np.uint8), "UMI")
    b.validate()
    return b, genome, contigs


def make_batch(cfg: FixedConfig, seed: int, n_pairs: Optional[int] = None, with_qnames: bool = True,
               genome_cache: Optional[Tuple[List[np.ndarray], Genome]] = None):
    """The batch of a BASELINE.json config: fixed-length shapes from make_fixed_batch, cfg5 from make_ragged_fixed."""
    if cfg.read_len_min > 0:
        return make_ragged_fixed(cfg, seed, n_pairs, with_qnames, genome_cache)
    return make_fixed_batch(cfg, seed, n_pairs, with_qnames, genome_cache)
