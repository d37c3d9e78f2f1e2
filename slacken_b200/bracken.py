"""Bracken weights (slacken/BrackenWeights.scala): all reads of a given length of the library's genomes, self-classified
against the library, counted by (destination taxon, source taxon), and written as a Bracken `kmer_distrib` file."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from ._lib import check
from .host import KeyValueIndex, pack_sequences

FRAGMENT_MAX = 1024 * 1024   # slacken/BrackenWeights.scala:302
TRIPLE_DTYPE = np.dtype([("dest", "<i4"), ("source", "<i4"), ("reads", "<u8")])


def split_to_max_length(seq: bytes, max_len: int, k: int) -> List[bytes]:
    """TaxonFragment.splitToMaxLength (slacken/BrackenWeights.scala:152-164); k is the READ length, so that every read lies
    in exactly one piece."""
    if len(seq) <= max_len:
        return [seq]
    return [seq[s:min(len(seq), s + max_len)] for s in range(0, len(seq) - k + 1, max_len - (k - 1))]


class BrackenWeights:
    def __init__(self, index: KeyValueIndex, read_len: int):
        self.index, self.read_len = index, read_len

    def build(self, genomes: Iterable[Tuple[int, bytes]], fragment_max: int = FRAGMENT_MAX,
              batch_bases: int = 1 << 30) -> Dict[Tuple[int, int], int]:
        """buildWeights (slacken/BrackenWeights.scala:312-354): {(dest, source): reads}. Every genome sequence is one
        TaxonFragment (whitespace-free), cut into pieces of at most fragment_max bases that overlap by read_len - 1."""
        out: Dict[Tuple[int, int], int] = {}
        pieces: List[bytes] = []
        taxa: List[int] = []
        size = 0

        def flush():
            nonlocal pieces, taxa, size
            if pieces:
                for d, s, c in self._run(pieces, taxa):
                    out[(d, s)] = out.get((d, s), 0) + c
            pieces, taxa, size = [], [], 0

        for taxon, seq in genomes:
            seq = seq.encode("latin-1") if isinstance(seq, str) else bytes(seq)
            for piece in split_to_max_length(seq, fragment_max, self.read_len):
                pieces.append(piece); taxa.append(int(taxon)); size += len(piece)
                if size >= batch_bases:
                    flush()
        flush()
        return out

    def _run(self, pieces: Sequence[bytes], taxa: Sequence[int]):
        ctx = self.index.ctx
        bases, off = pack_sequences(pieces)
        if len(bases) == 0:
            bases = np.zeros(16, dtype=np.uint8)
        ft = np.asarray(taxa, dtype=np.int32)
        cap = 64 * len(pieces)
        trip = np.zeros(cap, dtype=TRIPLE_DTYPE)
        n = C.c_uint64(0)
        check(ctx._L.slk_bracken_weights(self.index.h, bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                         ft.ctypes.data_as(C.c_void_p), len(pieces), self.read_len,
                                         trip.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        t = trip[:n.value]
        return zip(t["dest"].tolist(), t["source"].tolist(), t["reads"].tolist())


def kmer_distrib_lines(weights: Dict[Tuple[int, int], int]) -> List[str]:
    """writeKmerDistrib (slacken/BrackenWeights.scala:418-430). Spark leaves the order of lines and triples unspecified;
    here both are sorted."""
    total: Dict[int, int] = {}
    for (_, src), c in weights.items():
        total[src] = total.get(src, 0) + c
    by_dest: Dict[int, List[str]] = {}
    for (dest, src), c in sorted(weights.items()):
        by_dest.setdefault(dest, []).append(f"{src}:{c}:{total[src]}")
    return ["mapped_taxid\tgenome_taxids:kmers_mapped:total_genome_kmers"] + [f"{d}\t{' '.join(v)}" for d, v in sorted(by_dest.items())]


def write_kmer_distrib(weights: Dict[Tuple[int, int], int], path: str):
    with open(path, "w") as f:
        f.write("\n".join(kmer_distrib_lines(weights)) + "\n")
