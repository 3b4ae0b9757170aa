"""FASTA texts for the packing kernel (gcb_pack_fasta) and a restatement of the reference's reader to compare with.
Test infrastructure only."""
from __future__ import annotations

import numpy as np


def restate_fasta(text: bytes):
    """FastaReader's constructor + readNext + to4bits (fastareader.cpp:8-40, 58-102, 139-152) followed literally.
    Returns (ids, sizes, list of packed byte arrays) in file order."""
    n = len(text)
    pos = text.find(b">")          # fastareader.cpp:33-40: seek to the first '>'
    ids, sizes, packed = [], [], []
    if pos < 0:
        return ids, sizes, packed
    pos += 1
    while pos <= n:                 # readAll: while(!eof) readNext()
        if pos >= n and ids:        # the stream hit EOF inside the previous readNext
            break
        seq = bytearray()
        header = bytearray()
        found = False
        eof = False
        while True:
            if pos >= n:            # get(c) fails: eof
                eof = True
                break
            c = text[pos]
            pos += 1
            if c == ord(">"):
                break
            if found:
                if ord("a") <= c <= ord("z"):
                    c -= 32
                seq.append(c)       # the first char of a line is kept as it is (fastareader.cpp:74-78)
            else:
                header.append(c)
            end = text.find(b"\n", pos)
            line = text[pos:] if end < 0 else text[pos:end]
            pos = n if end < 0 else end + 1
            if end < 0:
                eof = True          # getline ran into EOF: the next get() fails
            if not found:
                header += line
                found = True
            else:                   # str_keep_valid_sequence (util.h:194-210)
                for ch in line:
                    if ord("a") <= ch <= ord("z"):
                        ch -= 32
                    if ord("A") <= ch <= ord("Z") or ch in (ord("-"), ord("*")):
                        seq.append(ch)
            if eof:
                break
        bits = np.zeros(len(seq), np.uint8)
        for ch, v in ((ord("A"), 1), (ord("T"), 2), (ord("C"), 3), (ord("G"), 4)):
            bits[np.frombuffer(bytes(seq), np.uint8) == ch] = v
        if len(bits) % 2:
            bits = np.concatenate([bits, np.zeros(1, np.uint8)])
        sp = bytes(header).find(b" ")
        ids.append(bytes(header) if sp < 0 else bytes(header[:sp]))
        sizes.append(len(seq))
        packed.append((bits[0::2] | (bits[1::2] << 4)).astype(np.uint8))
        if eof:
            break
    return ids, sizes, packed


def fasta_cases():
    rng = np.random.Generator(np.random.PCG64(11))

    def rnd(n, alphabet=b"ACGT"):
        return bytes(np.frombuffer(alphabet, np.uint8)[rng.integers(0, len(alphabet), n)])

    def wrap(seq, w):
        return b"\n".join(seq[i:i + w] for i in range(0, len(seq), w))

    cases = {
        "plain": b">chr1\n" + wrap(rnd(1000), 60) + b"\n>chr2 second contig\n" + wrap(rnd(333), 60) + b"\n",
        "no_trailing_newline": b">a\n" + wrap(rnd(125), 50),
        "junk_before_first_header": b"# comment\nACGT\n>x\nACGTACGTA\n",
        "header_mid_line": b"xx>y desc more\nACGT\nTT\n",
        "lower_case_and_iupac": b">m\nacgtnNRYKM\nacgu\n",
        "crlf": b">w one\r\nACGT\r\nGGCC\r\n>w2\r\nTT\r\n",
        "digits_dash_star_space": b">d\nAC 12-GT*\n9ACG T\n-A\n",
        "blank_lines": b">b\nACGT\n\nTTTT\nGG\n\n\nCC\n\n>swallowed\nAA\n>real\n\nC\n",
        "gt_inside_line": b">g\nAC>GT\nA>\n",
        "empty_contig": b">e1\n>e2\nACG\n>e3\n",
        "long_lines": b">L1\n" + rnd(9000) + b"\n>L2\n" + rnd(5000, b"ACGTN") + b"\n",
        "many_contigs": b"".join(b">c%d x\n" % i + wrap(rnd(int(rng.integers(1, 200))), 37) + b"\n" for i in range(40)),
        "only_header": b">h\n",
        "no_header": b"ACGT\nACGT\n",
        "random_bytes": b">r\n" + bytes(rng.integers(9, 127, 3000, dtype=np.uint8)).replace(b">", b"A") + b"\n",
        "block_boundaries": b">p\n" + wrap(rnd(3 * 4096 + 17), 4095) + b"\n>q\n" + rnd(4096 - 4) + b"\n\n" + rnd(10) + b"\n",
    }
    return cases


MALFORMED = {"gt_at_line_end": b">a\nAC\n>\nGG\n", "double_gt": b">>a\nAC\n", "gt_last_byte": b">a\nAC\n>"}
