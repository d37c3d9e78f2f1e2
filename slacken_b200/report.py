"""Kraken-style report and per-read output text (host side; the counts come from the device counters).

KrakenReport follows slacken/KrakenReport.scala:27-116, the per-read line follows ClassifiedRead.outputLine
(slacken/Classifier.scala:39-45) with TaxonCounts.pairsInOrderString / lengthString (slacken/TaxonCounts.scala:94-121)."""
from __future__ import annotations

import io
from decimal import ROUND_HALF_UP, Decimal
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

NONE, ROOT = 0, 1
_RANK_CODE = {"unclassified": "U", "root": "R", "superkingdom": "D", "kingdom": "K", "phylum": "P", "class": "C",
              "order": "O", "family": "F", "genus": "G", "species": "S"}  # slacken/Taxonomy.scala:38-47


def java_fixed(x: float, decimals: int, width: int = 0) -> str:
    """What java.util.Formatter prints for %<width>.<decimals>f: round HALF_UP on the decimal expansion of
    Double.toString(x) (the shortest representation), then pad on the left."""
    q = Decimal(1).scaleb(-decimals)
    return str(Decimal(repr(float(x))).quantize(q, rounding=ROUND_HALF_UP)).rjust(width)


class KrakenReport:
    def __init__(self, parents: np.ndarray, ranks: Sequence[Optional[str]], names: Sequence[Optional[str]],
                 counts: Iterable[Tuple[int, int]]):
        self.parents, self.ranks, self.names = parents, ranks, names
        counts = [(int(t), int(c)) for t, c in counts]
        self.taxon_counts = dict(counts)
        self.total = sum(c for _, c in counts)
        # TreeAggregator (slacken/KrakenReport.scala:27-41): every count flows to all ancestors
        self.clade_totals: dict = {}
        for t, c in counts:
            if t == NONE:
                self.clade_totals[NONE] = c
                continue
            x = t
            while x != NONE:
                self.clade_totals[x] = self.clade_totals.get(x, 0) + c
                x = int(parents[x])
        self._children: Optional[dict] = None

    def _kids(self, t: int):
        # only parents of taxa that were seen can have a non-zero child, so index lazily over the seen clades
        if self._children is None:
            ch: dict = {}
            for x in self.clade_totals:
                if x != NONE and x != ROOT:
                    ch.setdefault(int(self.parents[x]), []).append(x)
            # Taxonomy.children lists children in DESCENDING taxid order (prepend while scanning ascending ids,
            # slacken/Taxonomy.scala:193-201); the stable sort by clade count then keeps that order among ties
            for k in ch:
                ch[k].sort(reverse=True)
                ch[k].sort(key=lambda c: self.clade_totals.get(c, 0), reverse=True)
            self._children = ch
        return self._children.get(t, [])

    def _line(self, taxid: int, code: str, rank_depth: int, depth: int) -> str:
        clade, own = self.clade_totals.get(taxid, 0), self.taxon_counts.get(taxid, 0)
        perc = java_fixed(100.0 * clade / self.total, 2, 6)
        sub = "" if rank_depth == 0 else str(rank_depth)
        name = self.names[taxid] if self.names[taxid] is not None else ""
        return f"{perc}\t{clade}\t{own}\t{code}{sub}\t{taxid}\t{'  ' * depth}{name}"

    def text(self) -> str:
        out = io.StringIO()
        out.write("#Perc\tAggregate\tIn taxon\tRank\tTaxon\tName\n")
        if self.total == 0:
            return out.getvalue()
        if self.taxon_counts.get(NONE, 0) != 0:
            out.write(self._line(NONE, "U", 0, 0) + "\n")
        stack = [(ROOT, "R", 0, 0)]
        while stack:
            taxid, code, rank_depth, depth = stack.pop()
            title = "root" if taxid == ROOT else self.ranks[taxid]
            if title in _RANK_CODE:
                code, rank_depth = _RANK_CODE[title], 0
            else:
                rank_depth += 1
            out.write(self._line(taxid, code, rank_depth, depth) + "\n")
            for c in reversed(self._kids(taxid)):
                if self.clade_totals.get(c, 0) > 0:
                    stack.append((c, code, rank_depth, depth + 1))
        return out.getvalue()


def hits_string(hits) -> str:
    parts = []
    for t, c in zip(hits["taxon"].tolist(), hits["count"].tolist()):
        parts.append("|:|" if t == -2 else (f"A:{c}" if t == -1 else f"{t}:{c}"))
    return " ".join(parts)


def output_line(title: str, taxon: int, classified: bool, detail, hits) -> str:
    length = str(int(detail["len1"])) if int(detail["len2"]) == 0xFFFFFFFF else f"{int(detail['len1'])}|{int(detail['len2'])}"
    return f"{'C' if classified else 'U'}\t{title}\t{int(taxon)}\t{length}\t{hits_string(hits)}"
