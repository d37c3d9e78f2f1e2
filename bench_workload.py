"""The synthetic workload of BASELINE.json configs[1], shared by both arms of bench.py.
Pure numpy + seeds: no engine code, no oracle code."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RANK_TITLES = ["root", "superkingdom", "kingdom", "phylum", "class", "order", "family", "genus", "species"]


@dataclass
class Workload:
    name: str = "synthetic 10M x 150bp single-end reads vs 4 Gbp library (1000 genomes x 4 Mbp), k35/m31/s7, confidence 0.0"
    n_genomes: int = 1000
    genome_len: int = 4_000_000
    n_reads: int = 10_000_000
    read_len: int = 150
    tax_nodes: int = 50_000
    gseed: int = 1
    tseed: int = 2
    rseed: int = 3
    k: int = 35
    m: int = 31
    spaces: int = 7
    confidence: float = 0.0
    min_hit_groups: int = 2

    @property
    def total_bases(self) -> int:
        return self.n_genomes * self.genome_len


def taxonomy(w: Workload):
    """Tree shaped like the reference's test generator (src/test/scala/com/jnpersson/slacken/Testing.scala:62-83):
    equal node counts at the 8 ranks below root, each node's parent drawn from all shallower ids."""
    rng = np.random.default_rng(w.tseed)
    level = w.tax_nodes // 8 + 1
    n = 8 * level + 2
    parents = np.zeros(n, dtype=np.int32)
    ranks = [None] * n
    ranks[1] = "root"
    for d in range(1, 9):
        lo, hi = (d - 1) * level + 2, d * level + 2
        parents[lo:hi] = rng.integers(1, (d - 1) * level + 2, size=hi - lo)
        for t in range(lo, hi):
            ranks[t] = RANK_TITLES[d]
    parents[1] = 0
    names = [f"Taxon {t}" for t in range(n)]
    names[0] = "unclassified"
    species = np.arange(7 * level + 2, 8 * level + 2)
    genome_taxa = rng.choice(species, size=w.n_genomes, replace=w.n_genomes > len(species)).astype(np.int32)
    return parents, ranks, names, genome_taxa
