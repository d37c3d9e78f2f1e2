"""SURVEY section 8 row f1 on the GPU path: a library built on the GPU is written in Slacken's on-disk layout (bucketed snappy
Parquet `id1:int64, taxon:int32` + `.properties` + `_taxonomy/`, slacken/KeyValueIndex.scala:125-159), read back, loaded into
HBM (KeyValueIndex.loadRecords) and classifies like the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import Classifier, IndexParams, KeyValueIndex, Taxonomy
from slacken_b200 import library_io as lio
from slacken_b200.host import pack_sequences
from tests.test_gpu_parity import assert_batch_equal, make_world, oracle_lib
from tests.util import simulate_reads

pytestmark = pytest.mark.gpu


def _write_dmp(directory, parents, ranks, names):
    os.makedirs(directory)
    with open(os.path.join(directory, "nodes.dmp"), "w") as f:
        for t in range(1, len(parents)):
            if parents[t] != 0 or t == 1:
                f.write(f"{t}\t|\t{max(int(parents[t]), 1)}\t|\t{ranks[t] or 'no rank'}\t|\n")
    with open(os.path.join(directory, "names.dmp"), "w") as f:
        for t in range(1, len(parents)):
            if parents[t] != 0 or t == 1:
                f.write(f"{t}\t|\t{names[t]}\t|\t\t|\tscientific name\t|\n")


@pytest.mark.parametrize("k,m,s,buckets", [(35, 31, 7, 17), (28, 21, 3, 4)])
def test_gpu_built_library_round_trips_through_the_spark_layout(gpu, tmp_path, k, m, s, buckets):
    rng, parents, ranks, names, genomes, taxa = make_world(91)
    params = IndexParams(k=k, m=m, spaces=s, buckets=buckets)
    tax = Taxonomy(gpu, parents, ranks, names)
    pieces, labels = oracle.remove_invalid(genomes, taxa)      # InputReader.removeInvalid, as the Scala host does
    gb, go = pack_sequences(pieces)
    built = KeyValueIndex.build(gpu, tax, params, [(gb, go, np.asarray(labels, dtype=np.int32))], expected_bases=len(gb))
    id1, tx = built.records()
    dmp = str(tmp_path / "taxdump")
    _write_dmp(dmp, parents, ranks, names)
    loc = str(tmp_path / "lib" / "idx")
    os.makedirs(os.path.dirname(loc))
    lio.write_library(loc, params, id1, tx, taxonomy_dir=dmp)
    # the layout Spark expects: <loc>.properties, <loc>/part-*_<bucket:05d>.c000.snappy.parquet, <loc>_taxonomy/*.dmp
    assert os.path.exists(loc + ".properties") and os.path.exists(os.path.join(loc + "_taxonomy", "nodes.dmp"))
    files = sorted(glob.glob(os.path.join(loc, "part-*.snappy.parquet")))
    assert files and all(f.endswith(".c000.snappy.parquet") for f in files)
    for f in files[:3]:   # every row of a file hashes to the bucket in the file's name
        import pyarrow.parquet as pq
        bk = int(os.path.basename(f).split("_")[-1][:5])
        col = pq.read_table(f).column("id1").to_numpy()
        assert (lio.spark_bucket(col, buckets) == bk).all()
    p2, id2, tx2 = lio.read_library(loc)
    assert (p2.k, p2.m, p2.spaces, p2.canonical, p2.buckets) == (k, m, s, True, buckets)
    parents2, ranks2, names2 = lio.load_taxonomy_dmp(loc + "_taxonomy")
    tax2 = Taxonomy(gpu, parents2, ranks2, names2)
    loaded = KeyValueIndex.from_records(gpu, tax2, p2, id2, tx2)          # KeyValueIndex.loadRecords
    assert len(loaded) == len(built)
    a, b = loaded.records()
    assert np.array_equal(a, id1) and np.array_equal(b, tx)
    # ... and classifies like the oracle
    olib = oracle_lib(oracle.params(k=k, m=m, spaces=s), parents, genomes, taxa)
    oid, otx = olib.records()
    assert np.array_equal(oid, id1) and np.array_equal(otx, tx)
    reads = simulate_reads(rng, genomes, 1500, (20, 260), n_rate=0.1)
    rb, ro = pack_sequences(reads)
    cls = Classifier(loaded)
    got = cls.classify(rb, ro, confidence=0.1)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), confidence=0.1)
    assert_batch_equal(res, per, got, k)
    for o in (cls, loaded, built, tax2, tax):
        o.close()
