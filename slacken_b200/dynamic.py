"""Two-step classification with a sample-tailored ("dynamic") library: slacken/Dynamic.scala:250-374, host logic over the
same kernels. Step 1 finds the taxa present in the sample with one of three heuristics, step 2 rebuilds the library from
the genomes of those taxa (and all their descendants) and classifies the reads again.

    d = Dynamic(ctx, base_index, genomes, rank="species", criteria=ClassifiedReadCount(100, 0.15))
    taxon_set, dynamic_index = d.make_index(reads, "<out>_taxonSet.txt")
    Classifier(dynamic_index).classify(...)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from .host import Classifier, GpuContext, IndexParams, KeyValueIndex, ReportCounts, Taxonomy, pack_sequences

RANK_DEPTH = {"root": 0, "superkingdom": 1, "kingdom": 2, "phylum": 3, "class": 4, "order": 5, "family": 6, "genus": 7,
              "species": 8}   # slacken/Taxonomy.scala:38-47


@dataclass
class ClassifiedReadCount:   # --reads N (with the confidence of the first pass)
    threshold: int
    confidence: float = 0.0


@dataclass
class MinimizerTotalCount:   # --min-count N
    threshold: int


@dataclass
class MinimizerDistinctCount:   # --min-distinct N
    threshold: int


@dataclass
class GoldSetOptions:   # slacken/Dynamic.scala:55-62
    """taxon_file: one taxid per line (first CSV column) or an iterable of taxids; promote_rank: keep taxa that had to be
    promoted to an ancestor present in the library when they sit at this rank or below; classify_with: build the dynamic
    library from the gold set itself instead of only comparing the detected set with it."""
    taxon_file: object
    promote_rank: Optional[str] = None
    classify_with: bool = False


def format_perc(d: float) -> str:   # kmers/package.scala:60
    from .report import java_fixed
    return "NaN%" if d != d else java_fixed(d * 100, 2) + "%"


class TaxonomyTree:
    """The few tree queries of slacken/Taxonomy.scala that the taxon-set step needs."""

    def __init__(self, taxonomy: Taxonomy):
        self.parents, self.ranks = taxonomy.parents, taxonomy.ranks
        self._children: Optional[dict] = None

    def depth(self, t: int) -> int:   # slacken/Taxonomy.scala:222-228: rank depth of the nearest ranked node upwards
        while t != 0:
            r = self.ranks[t]
            if r in RANK_DEPTH:
                return RANK_DEPTH[r]
            t = int(self.parents[t])
        return -1

    def children(self, t: int) -> List[int]:
        if self._children is None:
            ch: dict = {}
            p = self.parents
            for x in np.nonzero(p)[0]:
                ch.setdefault(int(p[x]), []).append(int(x))
            self._children = ch
        return self._children.get(t, [])

    def with_descendants(self, taxa: Iterable[int]) -> Set[int]:   # slacken/Taxonomy.scala:314-320
        out, stack = set(int(t) for t in taxa), [int(t) for t in taxa]
        while stack:
            for c in self.children(stack.pop()):
                if c not in out:
                    out.add(c)
                    stack.append(c)
        return out

    def path_to_root(self, t: int):   # slacken/Taxonomy.scala:204-215
        t = int(t)
        while t != 0:
            yield t
            t = int(self.parents[t])

    def with_ancestors(self, taxa: Iterable[int]) -> Set[int]:   # slacken/Taxonomy.scala:307-311
        out: Set[int] = set()
        for a in taxa:
            for e in self.path_to_root(a):
                if e in out:
                    break
                out.add(e)
        return out

    def clade_totals(self, counts: Sequence[Tuple[int, int]]) -> dict:   # TreeAggregator, slacken/KrakenReport.scala:27-41
        tot: dict = {}
        for t, c in counts:
            x = int(t)
            while x != 0:
                tot[x] = tot.get(x, 0) + int(c)
                x = int(self.parents[x])
        return tot


def count_filter(tree: TaxonomyTree, counts: Sequence[Tuple[int, int]], rank: str, threshold: int) -> Set[int]:
    """CountFilter.taxa (slacken/Dynamic.scala:191-201): taxa with a count, at `rank` or below, whose clade total reaches
    the threshold."""
    tot = tree.clade_totals(counts)
    d = RANK_DEPTH[rank]
    return {int(t) for t, _ in counts if t != 0 and tree.depth(int(t)) >= d and tot.get(int(t), 0) >= threshold}


class Dynamic:
    def __init__(self, ctx: GpuContext, base: KeyValueIndex, genomes: Sequence[Tuple[int, bytes]], rank: str = "species",
                 criteria=ClassifiedReadCount(100, 0.0), min_hit_groups: int = 2,
                 gold_set_opts: Optional[GoldSetOptions] = None, primary: Optional[Sequence[int]] = None):
        """genomes: (taxon, sequence) pairs of the genome library the dynamic index is rebuilt from.
        primary: the merged.dmp mapping (secondary id -> primary id, slacken/Taxonomy.scala:100-103), identity if None."""
        self.ctx, self.base, self.genomes, self.rank, self.criteria = ctx, base, genomes, rank, criteria
        self.min_hit_groups = min_hit_groups
        self.gold_set_opts, self.primary = gold_set_opts, primary
        self.taxonomy = base.taxonomy
        self.tree = TaxonomyTree(self.taxonomy)
        self.log: List[str] = []   # what the reference prints to stdout, line by line
        self._in_library: Optional[Set[int]] = None

    def _say(self, msg: str):
        self.log.append(msg)

    # ---- the gold set (slacken/Dynamic.scala:282-318) ----------------------------------------------------------------
    def taxon_set_in_library(self) -> Set[int]:   # GenomeLibrary.taxonSet, slacken/GenomeLibrary.scala:35-44
        if self._in_library is None:
            self._in_library = self.tree.with_ancestors({int(t) for t, _ in self.genomes})
        return self._in_library

    def read_gold_set(self, opts: Optional[GoldSetOptions] = None) -> Set[int]:
        """readGoldSet: the taxa of the gold-set file (mapped through merged.dmp), those that have no sequence in the library
        promoted to their nearest ancestor that has, everything filtered at the reclassify rank; promoted taxa at
        promote_rank or below are kept whatever their depth."""
        opts = opts or self.gold_set_opts
        src = opts.taxon_file
        if isinstance(src, (str, bytes)):
            with open(src) as f:
                ids = [int(line.split(",")[0]) for line in f if line.strip()]
        else:
            ids = [int(x) for x in src]
        prim = (lambda x: x) if self.primary is None else (lambda x: int(self.primary[x]))
        gold = {prim(x) for x in ids}
        self._say(f"Gold set contained {len(gold)} taxa")
        lib = self.taxon_set_in_library()
        not_found = gold - lib
        promoted: Set[int] = set()
        for t in not_found:
            for a in self.tree.path_to_root(t):
                if a in lib:
                    promoted.add(a)
                    break
        self._say(f"{len(not_found)} taxa from gold set not found in library, promoted to {len(promoted)} taxa.")
        kept: Set[int] = set()
        if opts.promote_rank is not None:
            kept = {t for t in promoted if self.tree.depth(t) >= RANK_DEPTH[opts.promote_rank]}
            self._say(f"Keeping {len(kept)} taxa at rank {opts.promote_rank} and below from promoted set")
        total = gold | promoted
        filtered = {t for t in total if self.tree.depth(t) >= RANK_DEPTH[self.rank]} | kept
        self._say(f"Initial adjusted gold set size {len(total)}, filtered at {self.rank} to {len(filtered)}")
        return filtered

    def compare_with_gold_set(self, keep: Set[int]) -> dict:
        """The comparison findTaxonSet prints when a gold set is given (slacken/Dynamic.scala:265-275)."""
        gold = self.read_gold_set()
        tp = len(keep & gold)
        fp, fn = len(keep) - tp, len(gold) - tp
        precision = tp / (tp + fp) if tp + fp else float("nan")
        recall = tp / len(gold) if gold else float("nan")
        self._say(f"Comparing detected set with supplied gold set. True Positives: {tp}, False Positives: {fp}, "
                  f"False Negatives: {fn}, Precision: {format_perc(precision)}, Recall: {format_perc(recall)}")
        return {"tp": tp, "fp": fp, "fn": fn, "precision": precision, "recall": recall}

    # ---- the three counting methods (slacken/Dynamic.scala:84-145) -------------------------------------------------
    def classified_reads_per_taxon(self, bases1, off1, bases2=None, off2=None) -> List[Tuple[int, int]]:
        cls = Classifier(self.base)
        counts = ReportCounts(self.ctx, self.taxonomy, 1)
        cls.attach_counts(counts, 0)
        cls.classify(bases1, off1, bases2, off2, confidence=self.criteria.confidence, min_hit_groups=self.min_hit_groups,
                     per_read_output=False)
        v = counts.fetch(0)
        cls.attach_counts(None)
        counts.close(); cls.close()
        v[0] = 0   # only classified reads count (slacken/Dynamic.scala:137-143)
        return [(int(t), int(v[t])) for t in np.nonzero(v)[0]]

    def _span_hits(self, bases1, off1, bases2=None, off2=None):
        """(taxon, minimizer) of every sequence span whose minimizer has a record: findHitsWithMinimizers."""
        from .sharded import GpuSplitOps
        ops = GpuSplitOps(self.base, np.zeros(0, dtype=np.int32))
        try:
            n = len(off1) - 1
            d = [ops.upload(bases1 if len(bases1) else np.zeros(16, np.uint8)), ops.upload(off1.astype(np.uint64).view(np.int64))]
            d2 = [ops.upload(bases2), ops.upload(off2.astype(np.uint64).view(np.int64))] if bases2 is not None else [None, None]
            span_off, spans, n_spans = ops.scan_spans(d[0], d[1], d2[0], d2[1], n)
            keys, idx, _ = ops.route(spans, n_spans, 1)
            taxa = ops.probe(keys)
            k, t = keys.cpu().numpy().view(np.uint64), taxa.cpu().numpy()
            # k-mers of each span: the low 14 bits of its span word (slk_core.h: SLK_SPAN_CNT)
            c = (spans.cpu().numpy().view(np.uint64)[idx.cpu().numpy().astype(np.int64)] & np.uint64(0x3fff)).astype(np.int64)
        finally:
            ops.close()
        hit = t > 0
        keep = np.array([self.tree.depth(int(x)) >= RANK_DEPTH[self.rank] for x in np.unique(t[hit])], dtype=bool)
        ok_taxa = set(np.unique(t[hit])[keep].tolist())
        sel = hit & np.isin(t, list(ok_taxa))
        return t[sel], k[sel], c[sel]

    def total_minimizers_per_taxon(self, *reads) -> List[Tuple[int, int]]:
        t, _, _ = self._span_hits(*reads)
        u, c = np.unique(t, return_counts=True)
        return [(int(a), int(b)) for a, b in zip(u, c)]

    def distinct_minimizers_per_taxon(self, *reads) -> List[Tuple[int, int]]:
        t, k, _ = self._span_hits(*reads)
        pairs = np.unique(np.stack([t.astype(np.uint64), k]), axis=1)
        u, c = np.unique(pairs[0], return_counts=True)
        return [(int(a), int(b)) for a, b in zip(u, c)]

    # ---- IndexStatistics.showTaxonFullCoverageStats (slacken/IndexStatistics.scala:86-112) -----------------------------
    def minimizer_coverage(self, batch_bases: int = 256 << 20) -> dict:
        """For every taxon of the genome library: the super-mers of its genomes whose minimizer has a record in the base
        index, grouped by the depth of the record's (LCA) taxon. Returns {taxon: ([(depth, super-mers)], [(depth, distinct
        minimizers)])}, depths ascending. The genomes run through the span scan of the split path in batches (one genome
        = one "read"; a sequence span is a super-mer, MinSplitter.superkmerPositions = kmers/minimizer/MinSplitter.scala:180-216),
        the lookups through its probe; the (taxon, minimizer) counting is host work, as this is a report, not the hot path."""
        from .host import pack_sequences
        from .sharded import GpuSplitOps
        ops = GpuSplitOps(self.base, np.zeros(0, dtype=np.int32))
        parts = []   # per batch: unique (genome taxon, key) with the record's taxon and the number of super-mers
        try:
            at = 0
            while at < len(self.genomes):
                end, size = at, 0
                while end < len(self.genomes) and (end == at or size + len(self.genomes[end][1]) <= batch_bases):
                    size += len(self.genomes[end][1]); end += 1
                gt = np.array([int(t) for t, _ in self.genomes[at:end]], dtype=np.int64)
                bases, off = pack_sequences([g for _, g in self.genomes[at:end]])
                d_b = ops.upload(bases if len(bases) else np.zeros(16, np.uint8))
                d_o = ops.upload(off.astype(np.uint64).view(np.int64))
                span_off, spans, n_spans = ops.scan_spans(d_b, d_o, None, None, end - at)
                keys, idx, _ = ops.route(spans, n_spans, 1)
                taxa = ops.probe(keys).cpu().numpy().astype(np.int64)
                k = keys.cpu().numpy().view(np.uint64)
                so = span_off.cpu().numpy().astype(np.int64)
                genome_of = np.searchsorted(so, idx.cpu().numpy().astype(np.int64), side="right") - 1
                hit = taxa > 0                      # the join with the records is an inner join
                g, k, t = gt[genome_of[hit]], k[hit], taxa[hit]
                order = np.lexsort((k, g))
                g, k, t = g[order], k[order], t[order]
                head = np.ones(len(g), dtype=bool)
                head[1:] = (g[1:] != g[:-1]) | (k[1:] != k[:-1])
                starts = np.nonzero(head)[0]
                parts.append((g[head], k[head], t[head], np.diff(np.append(starts, len(g)))))
                at = end
        finally:
            ops.close()
        if not parts:
            return {}
        g, k, t, c = (np.concatenate([p[i] for p in parts]) for i in range(4))
        order = np.lexsort((k, g))                   # the same (taxon, minimizer) may occur in several batches
        g, k, t, c = g[order], k[order], t[order], c[order]
        head = np.ones(len(g), dtype=bool)
        head[1:] = (g[1:] != g[:-1]) | (k[1:] != k[:-1])
        count_all = np.add.reduceat(c, np.nonzero(head)[0]) if len(g) else c
        g, t = g[head], t[head]
        depth_of = {int(x): self.tree.depth(int(x)) for x in np.unique(t)}
        d = np.array([depth_of[int(x)] for x in t], dtype=np.int64)
        out = {}
        for taxon in np.unique(g):
            sel = g == taxon
            depths = np.unique(d[sel])
            out[int(taxon)] = ([(int(x), int(count_all[sel][d[sel] == x].sum())) for x in depths],
                               [(int(x), int((d[sel] == x).sum())) for x in depths])
        return out

    def write_minimizer_coverage(self, output_location: str, coverage: Optional[dict] = None) -> List[str]:
        """<out>_support_report_minimizerCoverage/ and <out>_support_report_minimizerDistinctCoverage/ (one text part file
        each; slacken/Dynamic.scala:230-243): "<taxon>  <depth>:<count>|<depth>:<count>..." per taxon of the genome library."""
        import os
        cov = self.minimizer_coverage() if coverage is None else coverage
        written = []
        for name, which in (("minimizerCoverage", 0), ("minimizerDistinctCoverage", 1)):
            d = f"{output_location}_support_report_{name}"
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "part-00000.txt"), "w") as f:
                for taxon in sorted(cov):
                    f.write(f"{taxon}  " + "|".join(f"{a}:{b}" for a, b in cov[taxon][which]) + "\n")
            open(os.path.join(d, "_SUCCESS"), "w").close()
            written.append(d)
        return written

    # ---- findTaxonSet + makeRecords (slacken/Dynamic.scala:250-280,362-374) --------------------------------------------
    def find_taxon_set(self, bases1, off1, bases2=None, off2=None, write_location: Optional[str] = None) -> Set[int]:
        c = self.criteria
        if isinstance(c, ClassifiedReadCount):
            counts = self.classified_reads_per_taxon(bases1, off1, bases2, off2)
        elif isinstance(c, MinimizerTotalCount):
            counts = self.total_minimizers_per_taxon(bases1, off1, bases2, off2)
        elif isinstance(c, MinimizerDistinctCount):
            counts = self.distinct_minimizers_per_taxon(bases1, off1, bases2, off2)
        else:
            raise ValueError("unknown taxon criteria")
        keep = count_filter(self.tree, counts, self.rank, c.threshold)
        if write_location:
            with open(write_location, "w") as f:   # the detected set BEFORE descendant expansion, one taxid per line
                for t in sorted(keep):
                    f.write(f"{t}\n")
        if self.gold_set_opts is not None:
            self.compare_with_gold_set(keep)
        expanded = self.tree.with_descendants(keep)
        self._say(f"Detected set: Initial scan (criterion {c}) produced {len(keep)} taxa at rank {self.rank}, "
                  f"expanded with descendants to {len(expanded)}")
        return expanded

    # ---- reportDynamicIndexSupport (slacken/Dynamic.scala:146-180,210-230): the four Kraken-style support reports ------
    def report_dynamic_index_support(self, output_location: str, bases1, off1, bases2=None, off2=None) -> List[str]:
        """<out>_support_report_{totalKmerCount,distinctMinimizerCount,totalMinimizerCount,classifiedReadCount}.txt:
        per taxon at the reclassify rank or below, the k-mers and (distinct) minimizers of the sample's sequence spans that
        hit it, and the reads classified to it at confidence 0; plus the two minimizerCoverage directories, which describe
        the genome library (write_minimizer_coverage)."""
        from .report import KrakenReport
        t, k, c = self._span_hits(bases1, off1, bases2, off2)
        taxa = np.unique(t)
        tot_kmers = [(int(a), int(c[t == a].sum())) for a in taxa]
        tot_mins = [(int(a), int((t == a).sum())) for a in taxa]
        dist_mins = [(int(a), int(len(np.unique(k[t == a])))) for a in taxa]
        saved = self.criteria
        self.criteria = ClassifiedReadCount(0, 0.0)   # initThreshold = 0.0 (slacken/Dynamic.scala:154)
        try:
            cls_reads = self.classified_reads_per_taxon(bases1, off1, bases2, off2)
        finally:
            self.criteria = saved
        tx = self.taxonomy
        written = []
        for name, counts in (("totalKmerCount", tot_kmers), ("distinctMinimizerCount", dist_mins),
                             ("totalMinimizerCount", tot_mins), ("classifiedReadCount", cls_reads)):
            path = f"{output_location}_support_report_{name}.txt"
            with open(path, "w") as f:
                f.write(KrakenReport(tx.parents, tx.ranks, tx.names, counts).text())
            written.append(path)
        written += self.write_minimizer_coverage(output_location)
        return written

    def make_index(self, bases1, off1, bases2=None, off2=None, write_location: Optional[str] = None,
                   gold_set: Optional[Iterable[int]] = None) -> Tuple[Set[int], KeyValueIndex]:
        """The dynamic library: records rebuilt from the genomes whose taxon is in the set
        (KeyValueIndex.makeRecords(library, Some(set)), slacken/KeyValueIndex.scala:102-116)."""
        if gold_set is None and self.gold_set_opts is not None and self.gold_set_opts.classify_with:
            gold_set = self.read_gold_set()   # makeRecords, slacken/Dynamic.scala:362-370
        taxon_set = (self.tree.with_descendants(gold_set) if gold_set is not None
                     else self.find_taxon_set(bases1, off1, bases2, off2, write_location))
        chosen = [(t, s) for t, s in self.genomes if int(t) in taxon_set]
        gb, goff = pack_sequences([s for _, s in chosen])
        taxa = np.array([t for t, _ in chosen], dtype=np.int32)
        index = KeyValueIndex.build(self.ctx, self.taxonomy, self.base.params, [(gb, goff, taxa)] if len(chosen) else [],
                                    expected_bases=len(gb))
        return taxon_set, index
