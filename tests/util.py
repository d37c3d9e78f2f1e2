"""Shared generators for the tests: random taxonomies shaped like the reference's scalacheck generator
(src/test/scala/com/jnpersson/slacken/Testing.scala:62-83), random genomes and simulated reads."""
from __future__ import annotations

import numpy as np

RANK_TITLES = ["root", "superkingdom", "kingdom", "phylum", "class", "order", "family", "genus", "species"]


def make_taxonomy(size: int, seed: int):
    """Equal number of nodes at each of the 8 ranks below root; the parent of a node is any node of a shallower
    level (Gen.choose(1, maxParent)). Returns (parents int32[n], ranks list, names list)."""
    rng = np.random.default_rng(seed)
    level = size // 8 + 1
    n = 8 * level + 2
    parents = np.zeros(n, dtype=np.int32)
    ranks = [None] * n
    names = [None] * n
    ranks[1], names[1] = "root", "Taxon 1"
    for d in range(1, 9):
        max_parent = (d - 1) * level + 1
        for t in range((d - 1) * level + 2, d * level + 2):
            parents[t] = int(rng.integers(1, max_parent + 1))
            ranks[t] = RANK_TITLES[d]
            names[t] = f"Taxon {t}"
    parents[1] = 0
    names[0] = "unclassified"
    return parents, ranks, names


def leaf_taxa(parents: np.ndarray):
    has_child = np.zeros(len(parents), dtype=bool)
    has_child[parents[parents > 0]] = True
    return [t for t in range(2, len(parents)) if parents[t] != 0 and not has_child[t]]


def random_dna(rng, n: int) -> bytes:
    return bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))


_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def simulate_reads(rng, genomes, n: int, length, sub_rate=0.01, n_rate=0.02, random_frac=0.2, lower_frac=0.05):
    """Reads drawn from the genomes (half reverse complemented, substitutions, occasional N runs, some lowercase,
    a few U), plus unrelated random reads. `length` is an int or a (lo, hi) range."""
    out = []
    for _ in range(n):
        L = length if isinstance(length, int) else int(rng.integers(length[0], length[1] + 1))
        if rng.random() < random_frac or not genomes:
            r = bytearray(random_dna(rng, L))
        else:
            g = genomes[int(rng.integers(len(genomes)))]
            if len(g) <= L:
                r = bytearray(g[:L])
            else:
                p = int(rng.integers(0, len(g) - L + 1))
                r = bytearray(g[p:p + L])
            if rng.random() < 0.5:
                r = bytearray(revcomp(bytes(r)))
            for i in range(len(r)):
                if rng.random() < sub_rate:
                    r[i] = b"ACGT"[int(rng.integers(4))]
        x = rng.random()
        if x < n_rate and len(r) > 0:
            p = int(rng.integers(len(r)))
            run = int(rng.choice([1, 1, 2, 5, 36, 40, 60]))
            for i in range(p, min(len(r), p + run)):
                r[i] = ord("N")
        if rng.random() < lower_frac:
            r = bytearray(bytes(r).lower())
        if rng.random() < 0.02:
            r = bytearray(bytes(r).replace(b"T", b"U", 3))
        out.append(bytes(r))
    return out


def chimeric_reads(rng, genomes, n: int, pieces: int, piece_len=(40, 90)):
    """Long reads stitched from many short genome slices: dozens of merged hits per read (exercises the paths for
    reads whose hit list outgrows the per-thread buffers)."""
    out = []
    for _ in range(n):
        parts = []
        for _ in range(pieces):
            g = genomes[int(rng.integers(len(genomes)))]
            L = int(rng.integers(piece_len[0], piece_len[1] + 1))
            p = int(rng.integers(0, len(g) - L))
            parts.append(g[p:p + L])
        out.append(b"".join(parts))
    return out
