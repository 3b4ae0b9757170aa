"""Minimal BAM writer/reader for the pipeline tests (BGZF over zlib, written from the SAM/BAM specification).
TEST INFRASTRUCTURE: turns a synthetic Batch into the coordinate-sorted BAM + FASTA a user would feed to gencore."""
from __future__ import annotations

import struct
import zlib
from typing import List, Tuple

import numpy as np

from gencore_b200.abi import Batch, Genome, align4

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    total = 18 + len(comp) + 8
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", total - 1) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bgzf(path: str, raw: bytes) -> None:
    with open(path, "wb") as f:
        for i in range(0, len(raw), 0xFF00):
            f.write(_bgzf_block(raw[i:i + 0xFF00]))
        f.write(BGZF_EOF)


def read_bgzf(path: str) -> bytes:
    out = []
    with open(path, "rb") as f:
        blob = f.read()
    i = 0
    while i < len(blob):
        xlen = struct.unpack_from("<H", blob, i + 10)[0]
        bsize = None
        j = i + 12
        while j < i + 12 + xlen:
            si1, si2, slen = blob[j], blob[j + 1], struct.unpack_from("<H", blob, j + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", blob, j + 4)[0]
            j += 4 + slen
        comp = blob[i + 12 + xlen:i + bsize + 1 - 8]
        out.append(zlib.decompress(comp, -15) if comp else b"")
        i += bsize + 1
    return b"".join(out)


def genome_to_fasta(path: str, contigs: List[np.ndarray], names: List[str]) -> None:
    with open(path, "w") as f:
        for name, c in zip(names, contigs):
            f.write(f">{name} synthetic\n")
            s = bytes(c).decode()
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")


def batch_to_bam(path: str, batch: Batch, genome: Genome, extras: bool = False) -> int:
    """Writes the reads of `batch` as a coordinate-sorted BAM.  Returns the number of records.
    extras: every 50th record is repeated as a secondary alignment, and five unmapped records end the file."""
    names = genome.names
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{int(l)}\n" for n, l in zip(names, genome.contig_len))
    hdr = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names))
    for n, l in zip(names, genome.contig_len):
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", int(l))
    cluster_of_pair = np.repeat(np.arange(batch.n_clusters), np.diff(batch.cluster_pair_off))
    recs: List[Tuple[int, int, int, bytes]] = []
    rd = batch.reads
    for p in range(batch.n_pairs):
        tid = int(batch.cluster_ref[cluster_of_pair[p]])
        qname = bytes(batch.qnames[p])
        for side in range(2):
            d = rd[2 * p + side]
            l = int(d["l_qseq"])
            if l < 0:
                continue
            m = rd[2 * p + 1 - side]
            has_mate = int(m["l_qseq"]) >= 0
            flag = (99 if side == 0 else 147) if has_mate else 73
            mtid, mpos = (tid, int(m["pos"])) if has_mate else (-1, -1)
            isize = int(d["isize"]) if has_mate else 0
            off = int(d["data_off"])
            qual = batch.payload[off:off + l].tobytes()
            seq = batch.payload[off + align4(l):off + align4(l) + (l + 1) // 2].tobytes()
            cig = batch.cigar[int(d["cigar_off"]):int(d["cigar_off"]) + int(d["n_cigar"])].astype("<u4").tobytes()
            aux = b"NMC" + bytes([int(batch.nm[2 * p + side])]) if batch.nm is not None else b""
            body = struct.pack("<iiBBHHHiiii", tid, int(d["pos"]), len(qname) + 1, 60, 4680, int(d["n_cigar"]), flag, l, mtid, mpos, isize)
            body += qname + b"\x00" + cig + seq + qual + aux
            recs.append((tid, int(d["pos"]), len(recs), struct.pack("<i", len(body)) + body))
    if extras:
        for k in range(0, len(recs), 50):
            tid, pos, _, rec = recs[k]
            body = bytearray(rec[4:])
            flag = struct.unpack_from("<H", body, 14)[0] | 0x100
            struct.pack_into("<H", body, 14, flag)
            recs.append((tid, pos, len(recs), rec[:4] + bytes(body)))
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    tail = b""
    if extras:
        for k in range(5):
            qn = b"unmapped%d" % k
            body = struct.pack("<iiBBHHHiiii", -1, -1, len(qn) + 1, 0, 4680, 0, 4, 4, -1, -1, 0) + qn + b"\x00" + bytes([0x12, 0x48]) + bytes([30] * 4)
            tail += struct.pack("<i", len(body)) + body
    write_bgzf(path, hdr + b"".join(r[3] for r in recs) + tail)
    return len(recs) + (5 if extras else 0)


def bam_records(path: str):
    """(header bytes, [(key, record bytes)]) with key = (tid, pos, mtid, mpos, isize), in file order."""
    raw = read_bgzf(path)
    l_text = struct.unpack_from("<i", raw, 4)[0]
    i = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, i)[0]
    i += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, i)[0]
        i += 4 + l_name + 4
    header = raw[:i]
    out = []
    while i < len(raw):
        block = struct.unpack_from("<i", raw, i)[0]
        rec = raw[i + 4:i + 4 + block]
        tid, pos = struct.unpack_from("<ii", rec, 0)
        mtid, mpos, isize = struct.unpack_from("<iii", rec, 20)
        out.append(((tid, pos, mtid, mpos, isize), rec))
        i += 4 + block
    return header, out


def assert_same_bam(path_a: str, path_b: str) -> int:
    """SURVEY §8c comparison rule: identical headers, identical records, order identical up to permutation inside runs
    tied on (tid, pos, mtid, mpos, isize) — the reference breaks those ties by heap address (gencore.h:35,41)."""
    ha, ra = bam_records(path_a)
    hb, rb = bam_records(path_b)
    assert ha == hb, "BAM headers differ"
    assert len(ra) == len(rb), f"record counts differ: {len(ra)} vs {len(rb)}"
    assert [k for k, _ in ra] == [k for k, _ in rb], "record keys / order differ"
    i = 0
    while i < len(ra):
        j = i
        while j < len(ra) and ra[j][0] == ra[i][0]:
            j += 1
        assert sorted(r for _, r in ra[i:j]) == sorted(r for _, r in rb[i:j]), f"records differ in the run at key {ra[i][0]}"
        i = j
    return len(ra)
