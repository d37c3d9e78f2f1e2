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


def java_double_to_string(x: float) -> str:
    """java.lang.Double.toString: the shortest digits that round-trip (as Python's repr finds them), printed in decimal
    notation for 1e-3 <= |x| < 1e7 and in "computerized scientific notation" (d.dddE[-]n, at least one digit after the point)
    otherwise."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    from decimal import Decimal
    d = Decimal(repr(x))
    sign, digits, exp = d.as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:   # normalise: no trailing zeros in the digit string
        digits.pop(); exp += 1
    sci_exp = exp + len(digits) - 1              # x = d.ddd * 10^sci_exp
    ds = "".join(str(v) for v in digits)
    neg = "-" if sign else ""
    if -3 <= sci_exp < 7:
        if sci_exp >= len(ds) - 1:
            return neg + ds + "0" * (sci_exp - (len(ds) - 1)) + ".0"
        if sci_exp >= 0:
            return neg + ds[:sci_exp + 1] + "." + ds[sci_exp + 1:]
        return neg + "0." + "0" * (-sci_exp - 1) + ds
    return neg + ds[0] + "." + (ds[1:] or "0") + "E" + str(sci_exp)


def threshold_string(threshold: float, thresholds: Sequence[float]) -> str:
    """`"%.Nf".format(threshold)` with N = the longest `num.toString.split("\\.")(1).length` among the thresholds
    (slacken/Classifier.scala:189-190): for 1.0E-5 that is the length of "0E-5"."""
    from .report import java_fixed
    return java_fixed(threshold, max(len(java_double_to_string(t).split(".")[1]) for t in thresholds))


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
