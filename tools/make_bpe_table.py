#!/usr/bin/env python
"""Derive the integer BPE merge table the engine ships from CLIP's published vocabulary.

Input : the gzip merges file the reference tokenizer loads
        (/root/reference/src/open_clip/tokenizer.py:26-28,144-146 - `bpe_simple_vocab_16e6.txt.gz`,
        lines [1 : 49152-256-2+1]).
Output: leaf_b200/data/clip_bpe_merges.bin - little-endian uint32[48894], entry r =
        (left_id << 16) | right_id for the merge of rank r. The merged symbol's id is 512 + r
        (tokenizer.py:147-153: vocab = 256 byte symbols, 256 byte symbols + '</w>', then the
        merges in file order). Nothing but integers is stored; the file is data, not code.

Run here (the build container, where /root/reference exists); the .bin is committed so the
GPU box never needs the reference tree.
"""
import gzip
import hashlib
import os
import struct
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/open_clip/bpe_simple_vocab_16e6.txt.gz"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "leaf_b200", "data", "clip_bpe_merges.bin")


def byte_symbol_order():
    """Order of the 256 byte symbols in the vocabulary (tokenizer.py:31-51)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return bs, [chr(c) for c in cs]


def main():
    raw = gzip.open(SRC).read().decode("utf-8").split("\n")
    merges = [tuple(m.split()) for m in raw[1:49152 - 256 - 2 + 1]]
    assert len(merges) == 48894
    bs, cs = byte_symbol_order()
    vocab = list(cs) + [c + "</w>" for c in cs]
    for m in merges:
        vocab.append("".join(m))
    enc = {}
    for i, v in enumerate(vocab):
        assert v not in enc, "vocabulary strings must be unique for id = 512 + rank to hold"
        enc[v] = i
    out = bytearray()
    for r, (a, b) in enumerate(merges):
        la, rb = enc[a], enc[b]
        assert enc[a + b] == 512 + r
        assert la < 65536 and rb < 65536
        out += struct.pack("<I", (la << 16) | rb)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    with open(DST, "wb") as f:
        f.write(out)
    print("wrote", os.path.normpath(DST), len(out), "bytes sha256", hashlib.sha256(out).hexdigest())


if __name__ == "__main__":
    main()
