"""Slacken's on-disk library layout, so a GPU-built library is readable by the Spark
driver and a Spark-built one loads into HBM:

    <idx>.properties          k, m, buckets, version=1, splitter=randomXOR, minimizerSpaces, XORmask, canonical
                              (kmers/IndexParams.scala:63-91, kmers/SplitterFormat.scala:55-77)
    <idx>/part-*_<bucket:05d>.c000.snappy.parquet   columns id1:int64, taxon:int32, one row per minimizer
                              (slacken/KeyValueIndex.scala:125-139; read side :150-159 trusts the bucket id in the name)
    <idx>_taxonomy/{nodes,names,merged}.dmp         (slacken/Taxonomy.scala:116-146)

Bucket id = pmod(Murmur3_x86_32.hashLong(id1, seed 42), buckets): Spark 3.5's HashPartitioning for a long column.
That arithmetic lives in Spark, not in the reference; no test of the reference pins it ("parity unpinned").
Host-side I/O only -- nothing here is on the measured path.
"""
from __future__ import annotations

import glob
import os
import shutil
import time
import uuid

import numpy as np

from .host import DEFAULT_TOGGLE_MASK, IndexParams

_M32 = np.uint32(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))) & _M32


def _mix_k1(k1):
    k1 = (k1 * np.uint32(0xCC9E2D51)) & _M32
    k1 = _rotl(k1, 15)
    return (k1 * np.uint32(0x1B873593)) & _M32


def _mix_h1(h1, k1):
    h1 = h1 ^ k1
    h1 = _rotl(h1, 13)
    return (h1 * np.uint32(5) + np.uint32(0xE6546B64)) & _M32


def spark_hash_long(values: np.ndarray, seed: int = 42) -> np.ndarray:
    """org.apache.spark.unsafe.hash.Murmur3_x86_32.hashLong, vectorised; returns int32."""
    v = np.ascontiguousarray(values).view(np.uint64)
    with np.errstate(over="ignore"):
        low = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        high = (v >> np.uint64(32)).astype(np.uint32)
        h1 = _mix_h1(np.full(v.shape, seed, dtype=np.uint32), _mix_k1(low))
        h1 = _mix_h1(h1, _mix_k1(high))
        h1 = h1 ^ np.uint32(8)
        h1 ^= h1 >> np.uint32(16)
        h1 = (h1 * np.uint32(0x85EBCA6B)) & _M32
        h1 ^= h1 >> np.uint32(13)
        h1 = (h1 * np.uint32(0xC2B2AE35)) & _M32
        h1 ^= h1 >> np.uint32(16)
    return h1.view(np.int32)


def spark_bucket(id1: np.ndarray, buckets: int) -> np.ndarray:
    return np.mod(spark_hash_long(id1).astype(np.int64), buckets).astype(np.int32)   # pmod


def write_properties(location: str, params: IndexParams, comment: str = "") -> None:
    """java.util.Properties.store layout: comment line, date line, key=value lines."""
    mask = params.toggle_mask if params.toggle_mask < (1 << 63) else params.toggle_mask - (1 << 64)
    props = {"k": params.k, "m": params.m, "buckets": params.buckets, "version": 1, "splitter": "randomXOR",
             "XORmask": mask, "canonical": "true" if params.canonical else "false"}
    if params.spaces:
        props["minimizerSpaces"] = params.spaces
    with open(location + ".properties", "w") as f:
        f.write(f"#{comment or 'Properties for Slacken KeyValueIndex ' + location}\n")
        f.write("#" + time.strftime("%a %b %d %H:%M:%S %Z %Y") + "\n")
        for k, v in props.items():
            f.write(f"{k}={v}\n")


def read_properties(location: str) -> IndexParams:
    props = {}
    for line in open(location + ".properties"):
        line = line.strip()
        if not line or line[0] in "#!":
            continue
        k, _, v = line.partition("=")
        props[k.strip()] = v.strip()
    if int(props.get("version", "1")) > 1:
        raise ValueError("a newer version of this software is needed to read " + location)   # IndexParams.scala:36-38
    if props.get("splitter", "standard") != "randomXOR":
        raise ValueError("only the randomXOR splitter is supported")
    mask = int(props["XORmask"]) & ((1 << 64) - 1) if "XORmask" in props else DEFAULT_TOGGLE_MASK
    return IndexParams(k=int(props["k"]), m=int(props["m"]), spaces=int(props.get("minimizerSpaces", "0")),
                       canonical=props.get("canonical", "true").lower() == "true", toggle_mask=mask,
                       buckets=int(props["buckets"]))


def write_records(location: str, id1: np.ndarray, taxon: np.ndarray, buckets: int) -> None:
    import pyarrow as pa
    import pyarrow.parquet as pq
    id1 = np.ascontiguousarray(id1).view(np.int64)
    taxon = np.ascontiguousarray(taxon, dtype=np.int32)
    if os.path.isdir(location):
        shutil.rmtree(location)     # SaveMode.Overwrite
    os.makedirs(location)
    b = spark_bucket(id1, buckets)
    order = np.argsort(b, kind="stable")
    bounds = np.searchsorted(b[order], np.arange(buckets + 1))
    uid = uuid.uuid4()
    schema = pa.schema([pa.field("id1", pa.int64(), nullable=False), pa.field("taxon", pa.int32(), nullable=False)])
    for bk in range(buckets):
        sel = order[bounds[bk]:bounds[bk + 1]]
        if len(sel) == 0:
            continue
        tbl = pa.Table.from_arrays([pa.array(id1[sel]), pa.array(taxon[sel])], schema=schema)
        pq.write_table(tbl, os.path.join(location, f"part-00000-{uid}_{bk:05d}.c000.snappy.parquet"), compression="snappy")
    open(os.path.join(location, "_SUCCESS"), "w").close()


def read_records(location: str):
    import pyarrow.parquet as pq
    ids, taxa = [], []
    for f in sorted(glob.glob(os.path.join(location, "*.parquet"))):
        t = pq.read_table(f, columns=["id1", "taxon"])
        ids.append(t.column("id1").to_numpy())
        taxa.append(t.column("taxon").to_numpy())
    if not ids:
        return np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.int32)
    return np.concatenate(ids).astype(np.int64).view(np.uint64), np.concatenate(taxa).astype(np.int32)


def write_library(location: str, params: IndexParams, id1: np.ndarray, taxon: np.ndarray, taxonomy_dir: str | None = None):
    """KeyValueIndex.writeRecords + Taxonomy.copyToLocation (slacken/Slacken.scala:160-163)."""
    write_properties(location, params)
    write_records(location, id1, taxon, params.buckets)
    if taxonomy_dir:
        os.makedirs(location + "_taxonomy", exist_ok=True)
        for name in ("nodes.dmp", "names.dmp", "merged.dmp"):
            src = os.path.join(taxonomy_dir, name)
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(location + "_taxonomy", name))


def read_library(location: str):
    """-> (IndexParams, id1 uint64[], taxon int32[]); the taxonomy is loaded with load_taxonomy_dmp."""
    return (read_properties(location),) + read_records(location)


def load_taxonomy_dmp(directory: str):
    """Taxonomy.load (slacken/Taxonomy.scala:116-137): -> (parents int32[], ranks, names). load_taxonomy_primary also
    returns merged.dmp's secondary -> primary mapping (Taxonomy.primary, slacken/Taxonomy.scala:100-103)."""
    return load_taxonomy_primary(directory)[:3]


def load_taxonomy_primary(directory: str):
    """-> (parents, ranks, names, primary int32[]): primary[t] = t unless merged.dmp maps the (secondary) id t to its primary id
    (Taxonomy.primary; Dynamic's gold set and the comparison tools apply it, slacken/Dynamic.scala:287)."""
    nodes = []
    for line in open(os.path.join(directory, "nodes.dmp")):
        x = line.split("|")
        nodes.append((int(x[0].strip()), int(x[1].strip()), x[2].strip()))
    merged = []
    mp = os.path.join(directory, "merged.dmp")
    if os.path.exists(mp):
        for line in open(mp):
            x = line.split("|")
            merged.append((int(x[0].strip()), int(x[1].strip())))
    n = max([t for t, _, _ in nodes] + [0]) + 1
    n = max(n, max([s for s, _ in merged] + [0]) + 1)
    parents = np.zeros(n, dtype=np.int32)
    ranks = [None] * n
    names = [None] * n
    for t, p, r in nodes:
        parents[t] = p
        ranks[t] = r
    for line in open(os.path.join(directory, "names.dmp")):
        x = line.split("|")
        if len(x) > 3 and x[3].strip() == "scientific name":
            t = int(x[0].strip())
            if 0 <= t < n:   # names of ids that nodes.dmp and merged.dmp do not know are ignored
                names[t] = x[1].strip()
    names[0] = "unclassified"
    parents[1] = 0
    ranks[0], ranks[1] = "unclassified", "root"
    primary = np.arange(n, dtype=np.int32)
    for sec, prim in merged:
        primary[sec] = prim
    return parents, ranks, names, primary
