"""Classification outputs on disk, laid out as Slacken writes them (slacken/Classifier.scala:184-251,417-422):

    <out>_c<thr>/sample=<sampleId>/part-00000.txt.gz     one line per read:  C|U \\t id \\t taxid \\t len|len1|len2 \\t hits
    <out>_c<thr>/<sampleId>_kreport.txt                  header line, optional unclassified line, DFS tree lines

<thr> is printed with as many decimals as the longest threshold given (slacken/Classifier.scala:189-191); sampleId is the
first group of --sample-regex in the read title, "other" without a match, "all" without a regex (slacken/Classifier.scala:
138-142). Reads without any span are neither written nor counted ("vanishing reads")."""
from __future__ import annotations

import gzip
import os
import re
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .host import ClassifiedBatch, Taxonomy
from .report import KrakenReport, output_line


def threshold_string(threshold: float, thresholds: Sequence[float]) -> str:
    """`"%.Nf".format(threshold)` with N = the longest decimal part among `thresholds` as printed by Double.toString."""
    def decimals(x: float) -> int:
        s = repr(float(x))
        return len(s.split(".")[1]) if "." in s and "e" not in s.lower() else 1
    from .report import java_fixed
    return java_fixed(threshold, max(decimals(t) for t in thresholds))


def sample_ids(titles: Sequence[str], sample_regex: Optional[str]) -> List[str]:
    if sample_regex is None:
        return ["all"] * len(titles)
    rx = re.compile(sample_regex)
    out = []
    for t in titles:
        m = rx.search(t)
        out.append(m.group(1) if m else "other")
    return out


class ClassificationWriter:
    """Accumulates classified batches per sample and writes the directory of one threshold."""

    def __init__(self, taxonomy: Taxonomy, output_location: str, threshold: float, thresholds: Sequence[float],
                 with_unclassified: bool = True, per_read_output: bool = True, sample_regex: Optional[str] = None):
        self.taxonomy = taxonomy
        self.location = output_location + "_c" + threshold_string(threshold, thresholds)
        self.with_unclassified, self.per_read_output, self.sample_regex = with_unclassified, per_read_output, sample_regex
        self.counts: Dict[str, Dict[int, int]] = {}
        self._files: Dict[str, gzip.GzipFile] = {}
        os.makedirs(self.location, exist_ok=True)

    def _file(self, sample: str):
        f = self._files.get(sample)
        if f is None:
            d = os.path.join(self.location, f"sample={sample}")
            os.makedirs(d, exist_ok=True)
            f = gzip.open(os.path.join(d, "part-00000.txt.gz"), "wt", encoding="utf-8", newline="\n")
            self._files[sample] = f
        return f

    def add(self, titles: Sequence[str], batch: ClassifiedBatch):
        samples = sample_ids(titles, self.sample_regex)
        keep = batch.has_span & (batch.classified | self.with_unclassified)
        for i in np.nonzero(keep)[0]:
            s = samples[i]
            c = self.counts.setdefault(s, {})
            t = int(batch.taxon[i])
            c[t] = c.get(t, 0) + 1
            if self.per_read_output:
                self._file(s).write(output_line(titles[i], t, bool(batch.classified[i]), batch.detail[i], batch.hits_of(i)) + "\n")

    def close(self) -> List[str]:
        """Writes <sample>_kreport.txt for every sample seen; returns the sample ids (slacken/Classifier.scala:245-251)."""
        for f in self._files.values():
            f.close()
        self._files = {}
        for s, c in self.counts.items():
            rep = KrakenReport(self.taxonomy.parents, self.taxonomy.ranks, self.taxonomy.names, sorted(c.items()))
            with open(os.path.join(self.location, f"{s}_kreport.txt"), "w", encoding="utf-8", newline="\n") as f:
                f.write(rep.text())
        return sorted(self.counts)
