"""ctypes/numpy mirror of include/gencore_b200.h (the C ABI of libgencore_b200.so).

Only layouts live here: no arithmetic.  `Batch` is the host-side container of one packed batch of
clusters — what a caller of Cluster::clusterByUMI (gencore.cpp:355,409) would hand over — and
`Result` the buffers the engine fills.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

GCB_ABI_VERSION = 2
GCB_MAX_UMI_WORDS = 4

GCB_OK, GCB_ERR_ARG, GCB_ERR_CUDA, GCB_ERR_NO_DEVICE, GCB_ERR_CAPACITY, GCB_ERR_MALFORMED = 0, -1, -2, -3, -4, -5

CLUSTER_CROSS_CONTIG = 0x01
CLUSTER_UMI_THR_SHIFT = 4

GROUP_DROPPED, GROUP_SSCS, GROUP_DCS, GROUP_DUPLEX_PARTNER, GROUP_DUPLEX_DIFF, GROUP_DUPLEX_SMALL = range(6)

STAGE_UMI_GROUP, STAGE_SELECT_TEMPLATE, STAGE_SCORE_VOTE, STAGE_DUPLEX, STAGE_ALL = 1, 2, 4, 8, 15
STAGE_VOTE_PREP_ONLY, STAGE_VOTE_ONLY = 16, 32  # measurement only: the two halves of STAGE_SCORE_VOTE
STAGE_VOTE_FAST_ONLY, STAGE_VOTE_REST_ONLY = 64, 128  # the two halves of STAGE_VOTE_ONLY: ring kernel; rollback + generic


def align4(x):
    return (x + 3) & ~3


class Options(C.Structure):
    """gcb_options; defaults are Options::Options (options.cpp:4-40)."""

    _fields_ = [
        ("duplex_mismatch_threshold", C.c_int32),
        ("cluster_size_req", C.c_int32),
        ("base_score_req", C.c_int32),
        ("high_quality", C.c_int32),
        ("moderate_quality", C.c_int32),
        ("low_quality", C.c_int32),
        ("score_high", C.c_int32),
        ("score_moderate", C.c_int32),
        ("score_low", C.c_int32),
        ("score_bad", C.c_int32),
        ("skip_low_complexity_cluster_threshold", C.c_int32),
        ("duplex_only", C.c_int32),
        ("disable_duplex", C.c_int32),
        ("reserved", C.c_int32),
        ("score_percent_req", C.c_double),
    ]

    @classmethod
    def default(cls, **kw) -> "Options":
        o = cls(
            duplex_mismatch_threshold=2, cluster_size_req=1, base_score_req=6, high_quality=30,
            moderate_quality=20, low_quality=15, score_high=8, score_moderate=6, score_low=4, score_bad=2,
            skip_low_complexity_cluster_threshold=1000, duplex_only=0, disable_duplex=0, reserved=0,
            score_percent_req=0.8,
        )
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        return o


READ_DESC = np.dtype(
    [
        ("data_off", "<i8"), ("l_qseq", "<i4"), ("pos", "<i4"), ("isize", "<i4"), ("cigar_off", "<i4"),
        ("n_cigar", "<u2"), ("l_qname", "<u2"), ("reserved", "<u4"),
    ],
    align=False,
)
assert READ_DESC.itemsize == 32

GROUP_RESULT = np.dtype(
    [
        ("out_off", "<i8", (2,)), ("tmpl_read", "<i4", (2,)), ("qname_donor", "<i4", (2,)),
        ("diff", "<i4", (2,)), ("mismatch_inc", "<i4", (2,)), ("merge_reads", "<i4"),
        ("reverse_merge_reads", "<i4"), ("status", "<i4"), ("duplex_partner", "<i4"),
        ("duplex_diff", "<i4"), ("umi_pair", "<i4"),
    ],
    align=False,
)
assert GROUP_RESULT.itemsize == 72


class BatchStruct(C.Structure):
    _fields_ = [
        ("n_clusters", C.c_int32), ("n_pairs", C.c_int32), ("umi_words", C.c_int32), ("max_cluster_bytes", C.c_int32),
        ("cluster_pair_off", C.c_void_p), ("cluster_ref", C.c_void_p), ("cluster_flags", C.c_void_p),
        ("umi", C.c_void_p), ("reads", C.c_void_p), ("cigar", C.c_void_p), ("n_cigar_ops", C.c_int64),
        ("payload", C.c_void_p), ("payload_bytes", C.c_int64),
    ]


class ResultStruct(C.Structure):
    _fields_ = [
        ("pair_group", C.c_void_p), ("cluster_n_groups", C.c_void_p), ("groups", C.c_void_p),
        ("out_payload", C.c_void_p), ("out_capacity", C.c_int64), ("out_bytes", C.c_void_p),
    ]


@dataclass
class Genome:
    """4-bit packed genome in the reference's own code (fastareader.cpp:139-152)."""

    packed4: np.ndarray      # uint8, all contigs back to back
    contig_off: np.ndarray   # int64 byte offset of each contig
    contig_len: np.ndarray   # int64 bases
    names: List[str] = field(default_factory=list)

    @staticmethod
    def from_bases(contigs: List[np.ndarray], names: Optional[List[str]] = None) -> "Genome":
        """contigs: uint8 arrays of ASCII bases."""
        lut = np.zeros(256, np.uint8)
        for ch, v in (("A", 1), ("T", 2), ("C", 3), ("G", 4)):
            lut[ord(ch)] = v
        parts, off, lens, cur = [], [], [], 0
        for c in contigs:
            bits = lut[c]
            if len(bits) % 2:
                bits = np.concatenate([bits, np.zeros(1, np.uint8)])
            packed = (bits[0::2] | (bits[1::2] << 4)).astype(np.uint8)
            pad = (-len(packed)) % 16
            if pad:
                packed = np.concatenate([packed, np.zeros(pad, np.uint8)])
            parts.append(packed)
            off.append(cur)
            lens.append(len(c))
            cur += len(packed)
        return Genome(
            np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(16, np.uint8),
            np.asarray(off, np.int64), np.asarray(lens, np.int64),
            names or [f"chr{i + 1}" for i in range(len(contigs))],
        )


@dataclass
class Batch:
    cluster_pair_off: np.ndarray  # int32 [n_clusters+1]
    cluster_ref: np.ndarray       # int32 [n_clusters]
    cluster_flags: np.ndarray     # uint8 [n_clusters]
    umi: np.ndarray               # uint64 [n_pairs, umi_words]
    reads: np.ndarray             # READ_DESC [2*n_pairs]
    cigar: np.ndarray             # uint32
    payload: np.ndarray           # uint8, len % 16 == 0
    # host-only companions (never cross the ABI): used by the reference harness and the host fix-ups
    qnames: Optional[List[bytes]] = None  # per pair, without NUL
    nm: Optional[np.ndarray] = None       # uint8 [2*n_pairs] NM:C per read
    umi_prefix: str = ""

    @property
    def n_clusters(self) -> int:
        return len(self.cluster_ref)

    @property
    def n_pairs(self) -> int:
        return len(self.umi)

    @property
    def umi_words(self) -> int:
        return self.umi.shape[1]

    def validate(self) -> None:
        assert self.cluster_pair_off.dtype == np.int32 and self.cluster_ref.dtype == np.int32
        assert self.cluster_flags.dtype == np.uint8 and self.umi.dtype == np.uint64
        assert self.reads.dtype == READ_DESC and self.cigar.dtype == np.uint32 and self.payload.dtype == np.uint8
        assert len(self.cluster_pair_off) == self.n_clusters + 1 and self.cluster_pair_off[-1] == self.n_pairs
        assert len(self.reads) == 2 * self.n_pairs and len(self.payload) % 16 == 0
        assert 1 <= self.umi_words <= GCB_MAX_UMI_WORDS
        for a in (self.cluster_pair_off, self.cluster_ref, self.cluster_flags, self.umi, self.reads, self.cigar, self.payload):
            assert a.flags["C_CONTIGUOUS"]

    def as_struct(self) -> BatchStruct:
        """Struct of HOST pointers (the arrays must stay alive while it is used)."""
        return BatchStruct(
            self.n_clusters, self.n_pairs, self.umi_words, self.max_cluster_bytes(),
            self.cluster_pair_off.ctypes.data, self.cluster_ref.ctypes.data, self.cluster_flags.ctypes.data,
            self.umi.ctypes.data, self.reads.ctypes.data, self.cigar.ctypes.data, len(self.cigar),
            self.payload.ctypes.data, len(self.payload),
        )

    def cluster_slab_bounds(self) -> np.ndarray:
        """int64 [n_clusters+1]: payload byte range of each cluster (see the header's layout rule)."""
        first = self.reads["data_off"][0::2]
        b = np.empty(self.n_clusters + 1, np.int64)
        b[:-1] = first[self.cluster_pair_off[:-1]]
        b[-1] = len(self.payload)
        return b

    def max_cluster_bytes(self) -> int:
        """gcb_batch.max_cluster_bytes: a property of the packed batch (whoever packs it knows it); computed once per object."""
        cached = self.__dict__.get("_max_cluster_bytes")
        if cached is None:
            cached = 0 if self.n_clusters == 0 else int(min(np.diff(self.cluster_slab_bounds()).max(), 2**31 - 1))
            self.__dict__["_max_cluster_bytes"] = cached
        return cached

    def algorithmic_bytes(self) -> dict:
        """SURVEY §8(d) payload-only byte counts for the vote kernel (input side)."""
        l = self.reads["l_qseq"].astype(np.int64)
        l = l[l >= 0]
        return {"reads_in": int(((l + 1) // 2 + l).sum())}


@dataclass
class Result:
    pair_group: np.ndarray        # int32 [n_pairs]
    cluster_n_groups: np.ndarray  # int32 [n_clusters]
    groups: np.ndarray            # GROUP_RESULT [n_pairs]
    out_payload: np.ndarray       # uint8
    out_bytes: np.ndarray         # int64 [1]

    @staticmethod
    def allocate(batch: Batch, out_capacity: Optional[int] = None) -> "Result":
        cap = len(batch.payload) if out_capacity is None else out_capacity
        return Result(
            np.full(batch.n_pairs, -1, np.int32), np.zeros(batch.n_clusters, np.int32),
            np.zeros(batch.n_pairs, GROUP_RESULT), np.zeros(max(cap, 16), np.uint8), np.zeros(1, np.int64),
        )

    def as_struct(self) -> ResultStruct:
        return ResultStruct(
            self.pair_group.ctypes.data, self.cluster_n_groups.ctypes.data, self.groups.ctypes.data,
            self.out_payload.ctypes.data, len(self.out_payload), self.out_bytes.ctypes.data,
        )

    def record(self, batch: Batch, slot: int, side: int):
        """(quals, packed bases) of the consensus record of group `slot`, side `side`, or None."""
        g = self.groups[slot]
        t = int(g["tmpl_read"][side])
        if t < 0:
            return None
        l = int(batch.reads["l_qseq"][t])
        off = int(g["out_off"][side])
        return self.out_payload[off:off + l], self.out_payload[off + align4(l):off + align4(l) + (l + 1) // 2]


UMI_CODE = {"A": 1, "C": 2, "G": 3, "T": 4, "_": 5}


def encode_umi(umi: str, words: int) -> np.ndarray:
    """String -> 4-bit fields (see the header's encoding conventions)."""
    if len(umi) > 16 * words:
        raise ValueError(f"UMI {umi!r} longer than {16 * words}")
    out = np.zeros(words, np.uint64)
    for k, ch in enumerate(umi):
        out[k >> 4] |= np.uint64(UMI_CODE[ch] << (60 - 4 * (k & 15)))
    return out


def padded_l_qname(name_len: int) -> int:
    """core.l_qname as htslib stores it in memory: name + NUL, padded to a multiple of 4 (Q14)."""
    return (name_len + 1 + 3) & ~3
