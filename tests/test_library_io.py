"""The Slacken library layout on disk (properties + bucketed snappy Parquet + _taxonomy)."""
import glob
import os

import numpy as np

from slacken_b200 import IndexParams
from slacken_b200 import library_io as lio


def java_murmur3_hash_long(v: int, seed: int = 42) -> int:
    """Scalar transcription of Murmur3_x86_32.hashLong, for cross-checking the vectorised version."""
    M = 0xFFFFFFFF

    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & M

    def mix_k1(k):
        k = (k * 0xCC9E2D51) & M
        k = rotl(k, 15)
        return (k * 0x1B873593) & M

    def mix_h1(h, k):
        h ^= k
        h = rotl(h, 13)
        return (h * 5 + 0xE6546B64) & M
    v &= (1 << 64) - 1
    h = mix_h1(seed, mix_k1(v & M))
    h = mix_h1(h, mix_k1(v >> 32))
    h ^= 8
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M
    h ^= h >> 16
    return h - (1 << 32) if h >= (1 << 31) else h


def test_spark_bucket_hash():
    rng = np.random.default_rng(0)
    vals = rng.integers(0, 2**63, 200, dtype=np.int64)
    vals[:4] = [0, 1, -1, 42]
    h = lio.spark_hash_long(vals)
    assert [int(x) for x in h] == [java_murmur3_hash_long(int(v)) for v in vals]
    b = lio.spark_bucket(vals, 200)
    assert b.min() >= 0 and b.max() < 200


def test_library_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    id1 = np.unique(rng.integers(0, 2**63, 5000, dtype=np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFFCCCCCCCC))
    taxon = rng.integers(1, 1000, len(id1)).astype(np.int32)
    tdir = tmp_path / "tax"
    tdir.mkdir()
    (tdir / "nodes.dmp").write_text("1\t|\t1\t|\tno rank\t|\n2\t|\t1\t|\tsuperkingdom\t|\n9\t|\t2\t|\tspecies\t|\n")
    (tdir / "names.dmp").write_text("1\t|\troot\t|\t\t|\tscientific name\t|\n9\t|\tBug\t|\t\t|\tscientific name\t|\n9\t|\tbuggy\t|\t\t|\tsynonym\t|\n")
    (tdir / "merged.dmp").write_text("12\t|\t9\t|\n")
    loc = str(tmp_path / "lib" / "idx")
    os.makedirs(os.path.dirname(loc))
    params = IndexParams(buckets=16)
    lio.write_library(loc, params, id1, taxon, str(tdir))
    text = open(loc + ".properties").read()
    assert "XORmask=-2054159557099562451" in text and "splitter=randomXOR" in text and "minimizerSpaces=7" in text
    p2, id2, tx2 = lio.read_library(loc)
    assert (p2.k, p2.m, p2.spaces, p2.canonical, p2.toggle_mask, p2.buckets) == (35, 31, 7, True, params.toggle_mask, 16)
    o1, o2 = np.argsort(id1), np.argsort(id2)
    assert np.array_equal(id1[o1], id2[o2]) and np.array_equal(taxon[o1], tx2[o2])
    import pyarrow.parquet as pq
    for f in glob.glob(os.path.join(loc, "*.parquet")):     # every row sits in the bucket its file name claims
        bk = int(f.split("_")[-1].split(".")[0])
        ids = pq.read_table(f).column("id1").to_numpy()
        assert (lio.spark_bucket(ids, 16) == bk).all()
    parents, ranks, names = lio.load_taxonomy_dmp(loc + "_taxonomy")
    assert len(parents) == 13 and parents[9] == 2 and parents[1] == 0 and names[9] == "Bug" and ranks[9] == "species"
