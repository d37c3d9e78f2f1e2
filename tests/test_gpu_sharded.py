"""The split (sharded-library) path on the GPU through the C ABI: scan -> route -> probe on the owner -> resolve.
One GPU plays every rank: the library is cut into `world` shards by slk_shard_of_records, the keys routed by the
device kernel are handed to the shard that owns them, and the results must equal the oracle that holds the whole
library (and therefore the fused classify kernel). The NCCL exchange itself is exercised by bench_sharded.py."""
import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import IndexParams, KeyValueIndex, Taxonomy
from slacken_b200.host import pack_sequences
from slacken_b200.sharded import GpuSplitOps, ShardedClassifier, ShardedKeyValueIndex, shard_of_records
from tests.test_gpu_parity import assert_batch_equal, make_world, oracle_lib
from tests.util import simulate_reads

pytestmark = pytest.mark.gpu


def _world(gpu, seed):
    rng, parents, ranks, names, genomes, taxa = make_world(seed)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    return rng, genomes, olib, id1, tx, tax


@pytest.mark.parametrize("paired", [False, True])
def test_split_path_one_shard_equals_oracle(gpu, paired):
    rng, genomes, olib, id1, tx, tax = _world(gpu, 31)
    shard = ShardedKeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx, rank=0, world=1)
    assert len(shard) == len(id1)
    cls = ShardedClassifier(shard)
    reads = simulate_reads(rng, genomes, 1500, (10, 260), n_rate=0.15) + [b"", b"ACGT", b"N" * 80]
    mates = simulate_reads(rng, genomes, len(reads), (10, 260), n_rate=0.15) if paired else None
    rb, ro = pack_sequences(reads)
    mb, mo = pack_sequences(mates) if paired else (None, None)
    for conf in (0.0, 0.2):
        got = cls.classify(rb, ro, mb, mo, confidence=conf)
        res, _, _, per = olib.classify(rb, ro.astype(np.int64), mb, mo.astype(np.int64) if paired else None, confidence=conf)
        assert_batch_equal(res, per, got, 35)
    cls.close(); shard.close(); tax.close()


@pytest.mark.parametrize("world", [2, 5])
def test_split_path_many_shards_on_one_gpu(gpu, world):
    rng, genomes, olib, id1, tx, tax = _world(gpu, 37)
    params = IndexParams()
    owner = shard_of_records(params, id1, world)
    shards = [KeyValueIndex.from_records(gpu, tax, params, id1[owner == r], tx[owner == r], world=world) for r in range(world)]
    assert sum(len(s) for s in shards) == len(id1) and all(len(s) > 0 for s in shards)
    union = np.unique(tx)
    ops = [GpuSplitOps(s, union) for s in shards]
    reads = simulate_reads(rng, genomes, 2000, (30, 200), n_rate=0.1)
    rb, ro = pack_sequences(reads)
    q = ops[0]   # rank 0 asks, every shard answers
    d_b, d_o = q.upload(rb), q.upload(ro.view(np.int64))
    span_off, spans, n_spans = q.scan_spans(d_b, d_o, None, None, len(reads))
    keys, idx, counts = q.route(spans, n_spans, world)
    assert sum(counts) == int(keys.numel()) and all(c > 0 for c in counts)
    import torch
    starts = np.concatenate([[0], np.cumsum(counts)])
    taxa = torch.cat([ops[r].probe(keys[starts[r]:starts[r + 1]].contiguous()) for r in range(world)])   # the "all-to-all"
    got = q.resolve(spans, span_off, n_spans, len(reads), False, idx, taxa, 0.1, 2, True)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), confidence=0.1)
    assert_batch_equal(res, per, got, 35)
    # a key asked of the wrong shard is a miss there: the owner function of the host and of the device agree
    wrong = ops[1].probe(keys[starts[0]:starts[1]].contiguous())
    assert int((wrong != 0).sum()) == 0
    for o in ops:
        o.close()
    for s in shards:
        s.close()
    tax.close()


def _mailbox_world(devices, seed=41, n_reads=1800, cap=120000):
    """One process plays every rank of the NVLink-mailbox exchange: rank r lives on devices[r] (several ranks may share a
    device), holds shard r and classifies its own share of the reads."""
    from slacken_b200.host import GpuContext
    from slacken_b200.sharded import Mailbox
    world = len(devices)
    rng, parents, ranks, names, genomes, taxa = make_world(seed)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    params = IndexParams()
    owner = shard_of_records(params, id1, world)
    ctxs = {d: GpuContext(d) for d in sorted(set(devices))}
    taxs = {d: Taxonomy(c, parents, ranks, names) for d, c in ctxs.items()}
    union = np.unique(tx)
    cls = []
    for r, d in enumerate(devices):
        shard = ShardedKeyValueIndex(KeyValueIndex.from_records(ctxs[d], taxs[d], params, id1[owner == r], tx[owner == r], world=world), r, world)
        cls.append(ShardedClassifier(shard, mailbox=Mailbox(ctxs[d], r, world, cap, connect=False), taxa_union=union))
    Mailbox.connect_local([c.mailbox for c in cls])
    return rng, genomes, olib, cls, (ctxs, taxs)


def _mailbox_round(cls, olib, per_rank_reads, conf, paired_mates=None):
    """scan on every rank, then route, probe, resolve: the order one-process emulation needs so that no kernel waits for a
    signal that is queued behind it on the same stream."""
    world = len(cls)
    st = []
    for r in range(world):
        rb, ro = pack_sequences(per_rank_reads[r])
        ops = cls[r].ops
        d_b, d_o = ops.upload(rb if len(rb) else np.zeros(16, np.uint8)), ops.upload(ro.view(np.int64))
        d_b2 = d_o2 = mb = mo = None
        if paired_mates is not None:
            mb, mo = pack_sequences(paired_mates[r])
            d_b2, d_o2 = ops.upload(mb if len(mb) else np.zeros(16, np.uint8)), ops.upload(mo.view(np.int64))
        span_off, spans, n_spans = ops.scan_spans(d_b, d_o, d_b2, d_o2, len(per_rank_reads[r]))
        st.append((rb, ro, mb, mo, span_off, spans, n_spans))
    for r in range(world):
        cls[r].mailbox_route(st[r][5], st[r][6])
    for r in range(world):
        cls[r].mailbox_probe()
    for r in range(world):
        rb, ro, mb, mo, span_off, spans, n_spans = st[r]
        n = len(per_rank_reads[r])
        got = cls[r].mailbox_resolve(spans, span_off, n_spans, n, paired_mates is not None, conf, 2, True)
        if n:
            res, _, _, per = olib.classify(rb, ro.astype(np.int64), mb, mo.astype(np.int64) if mo is not None else None, confidence=conf)
            assert_batch_equal(res, per, got, 35)


@pytest.mark.parametrize("world", [1, 2, 5])
def test_mailbox_exchange_ranks_on_one_gpu(gpu, world):
    rng, genomes, olib, cls, keep = _mailbox_world([gpu.device] * world)
    reads = simulate_reads(rng, genomes, 1800, (30, 220), n_rate=0.1) + [b"", b"ACGT", b"N" * 80]
    share = [reads[r::world] for r in range(world)]
    _mailbox_round(cls, olib, share, 0.0)
    # a second batch through the same mailboxes (next epoch), one rank without reads, paired, other confidence
    reads2 = simulate_reads(rng, genomes, 900, (30, 220), n_rate=0.1)
    mates2 = simulate_reads(rng, genomes, 900, (30, 220), n_rate=0.1)
    share2 = [reads2[r::world] for r in range(world)]
    mshare2 = [mates2[r::world] for r in range(world)]
    if world > 1:
        share2[world - 1], mshare2[world - 1] = [], []
    _mailbox_round(cls, olib, share2, 0.15, mshare2)
    for c in cls:
        c.close()


def test_mailbox_empty_batch_then_full_batch_and_teardown(gpu):
    """A rank that sends nothing to any owner still consumes every owner's completion flag (so the next batch may reuse the
    flags and buffers), several epochs in a row, and the mailboxes can be torn down right after the last batch."""
    world = 3
    rng, genomes, olib, cls, keep = _mailbox_world([gpu.device] * world)
    reads = simulate_reads(rng, genomes, 900, (30, 220), n_rate=0.1)
    for rnd in range(4):
        share = [reads[r::world] for r in range(world)]
        share[rnd % world] = []            # this rank has no reads at all in this batch
        share[(rnd + 1) % world] = [b"N" * 90, b"ACG"]   # ... and this one only reads without a sequence span
        _mailbox_round(cls, olib, share, 0.0)
    _mailbox_round(cls, olib, [reads[r::world] for r in range(world)], 0.15)
    for c in cls:
        c.close()


def test_mailbox_overflow_fails_loudly(gpu):
    from slacken_b200._lib import SLK_E_NOSPACE, SlackenGpuError
    rng, genomes, olib, cls, keep = _mailbox_world([gpu.device] * 2, cap=64)
    reads = simulate_reads(rng, genomes, 400, (100, 200))
    with pytest.raises(SlackenGpuError) as e:
        _mailbox_round(cls, olib, [reads[0::2], reads[1::2]], 0.0)
    assert e.value.code == SLK_E_NOSPACE
    for c in cls:
        c.close()


def test_mailbox_exchange_over_nvlink_two_gpus(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng, genomes, olib, cls, keep = _mailbox_world([0, 1, 0, 1], n_reads=3000)
    reads = simulate_reads(rng, genomes, 3000, (30, 220), n_rate=0.1)
    for conf in (0.0, 0.1):
        _mailbox_round(cls, olib, [reads[r::4] for r in range(4)], conf)
    for c in cls:
        c.close()


def test_records_stay_on_the_device(gpu):
    """The pieces of the distributed build that keep the records in HBM: table -> device rows, owner of every row on the
    device (equal to the host's), device rows -> table."""
    import ctypes as C

    import torch
    from slacken_b200._lib import check
    rng, genomes, olib, id1, tx, tax = _world(gpu, 43)
    params = IndexParams()
    full = KeyValueIndex.from_records(gpu, tax, params, id1, tx)
    d_id, d_tx = full.records_dev()
    assert d_id.is_cuda and d_id.numel() == len(id1)
    o = np.argsort(d_id.cpu().numpy().view(np.uint64), kind="stable")
    assert np.array_equal(d_id.cpu().numpy().view(np.uint64)[o], id1) and np.array_equal(d_tx.cpu().numpy()[o], tx)
    owner = torch.empty(d_id.numel(), dtype=torch.uint8, device=d_id.device)
    p = params.c_params()
    check(gpu._L.slk_shard_of_records_dev(gpu.h, C.byref(p), C.c_void_p(d_id.data_ptr()), d_id.numel(), 3, C.c_void_p(owner.data_ptr())))
    assert np.array_equal(owner.cpu().numpy(), shard_of_records(params, d_id.cpu().numpy(), 3))
    mine = owner == 1
    part = KeyValueIndex.from_records_dev(gpu, tax, params, d_id[mine], d_tx[mine], world=3)
    pid, ptx = part.records()
    sel = shard_of_records(params, id1, 3) == 1
    assert np.array_equal(pid, id1[sel]) and np.array_equal(ptx, tx[sel])
    # the export grouped by owner: the rows of owner d are a contiguous range, and each range holds exactly d's records
    g_id = torch.empty(len(id1), dtype=torch.int64, device=d_id.device)
    g_tx = torch.empty(len(id1), dtype=torch.int32, device=d_id.device)
    cnt = (C.c_uint64 * 3)()
    check(gpu._L.slk_index_records_by_owner_dev(full.h, 3, C.c_void_p(g_id.data_ptr()), C.c_void_p(g_tx.data_ptr()), len(id1), cnt))
    host_owner = shard_of_records(params, id1, 3)
    assert [int(c) for c in cnt] == np.bincount(host_owner, minlength=3).tolist()
    at = 0
    for d in range(3):
        rows_id = g_id[at:at + int(cnt[d])].cpu().numpy().view(np.uint64)
        rows_tx = g_tx[at:at + int(cnt[d])].cpu().numpy()
        o2 = np.argsort(rows_id, kind="stable")
        assert np.array_equal(rows_id[o2], id1[host_owner == d]) and np.array_equal(rows_tx[o2], tx[host_owner == d])
        at += int(cnt[d])
    part.close(); full.close(); tax.close()


def test_pipelined_batches_through_the_mailbox(gpu):
    """classify_pipelined: the scan of the next batch overlaps the exchange of the current one; results per batch must
    still equal the oracle's (one rank, three batches of different sizes, one of them paired, one empty)."""
    from slacken_b200.sharded import Mailbox
    rng, genomes, olib, id1, tx, tax = _world(gpu, 47)
    shard = ShardedKeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx, rank=0, world=1)
    cls = ShardedClassifier(shard, mailbox=Mailbox(gpu, 0, 1, 200000), taxa_union=np.unique(tx))
    sets = [simulate_reads(rng, genomes, 1200, (30, 220), n_rate=0.1), simulate_reads(rng, genomes, 700, (30, 220)), [],
            simulate_reads(rng, genomes, 900, (30, 220), n_rate=0.05)]
    mates = [None, simulate_reads(rng, genomes, 700, (30, 220)), None, None]
    host, dev = [], []
    for reads, m in zip(sets, mates):
        rb, ro = pack_sequences(reads)
        mb, mo = pack_sequences(m) if m is not None else (None, None)
        host.append((rb, ro, mb, mo))
        up = cls.ops.upload
        dev.append((up(rb if len(rb) else np.zeros(16, np.uint8)), up(ro.view(np.int64)),
                    up(mb) if m is not None else None, up(mo.view(np.int64)) if m is not None else None, len(reads)))
    n_out = 0
    for (rb, ro, mb, mo), got in zip(host, cls.classify_pipelined(dev, confidence=0.1)):
        if len(ro) > 1:
            res, _, _, per = olib.classify(rb, ro.astype(np.int64), mb, mo.astype(np.int64) if mo is not None else None, confidence=0.1)
            assert_batch_equal(res, per, got, 35)
        n_out += 1
    assert n_out == len(sets)
    cls.close(); shard.close(); tax.close()


def test_span_scan_one_pass_and_two_pass_agree(gpu, monkeypatch):
    """slk_scan_spans_dev + slk_emit_spans_dev: the one-pass scan (rows of a scratch, then compaction), the two-pass scan
    forced by the environment, and the two-pass scan a very long read falls back to give the same span words."""
    from tests.util import chimeric_reads
    rng, genomes, olib, id1, tx, tax = _world(gpu, 59)
    shard = ShardedKeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx, rank=0, world=1)
    ops = GpuSplitOps(shard.index, np.unique(tx))
    reads = simulate_reads(rng, genomes, 1500, (10, 300), n_rate=0.15) + [b"", b"ACGT", b"N" * 90]
    mates = simulate_reads(rng, genomes, len(reads), (10, 300), n_rate=0.15)
    rb, ro = pack_sequences(reads)
    mb, mo = pack_sequences(mates)
    d = [ops.upload(rb), ops.upload(ro.view(np.int64)), ops.upload(mb), ops.upload(mo.view(np.int64))]
    for paired in (False, True):
        args = (d[0], d[1], d[2] if paired else None, d[3] if paired else None, len(reads))
        monkeypatch.delenv("SLK_SPANS_TWO_PASS", raising=False)
        off1, spans1, n1 = ops.scan_spans(*args)
        monkeypatch.setenv("SLK_SPANS_TWO_PASS", "1")
        off2, spans2, n2 = ops.scan_spans(*args)
        assert n1 == n2 and n1 > 10000
        assert np.array_equal(off1.cpu().numpy(), off2.cpu().numpy())
        assert np.array_equal(spans1.cpu().numpy()[:n1], spans2.cpu().numpy()[:n2])
    monkeypatch.delenv("SLK_SPANS_TWO_PASS", raising=False)
    # one read of 6000 bases makes the rows too long for the scratch: the call falls back to two passes by itself
    long_reads = reads[:200] + chimeric_reads(rng, genomes, 1, 80, (70, 80))
    assert max(len(r) for r in long_reads) > 4200
    lb, lo = pack_sequences(long_reads)
    cls = ShardedClassifier(shard)
    got = cls.classify(lb, lo, confidence=0.05)
    res, _, _, per = olib.classify(lb, lo.astype(np.int64), confidence=0.05)
    assert_batch_equal(res, per, got, 35)
    ops.close(); cls.close(); shard.close(); tax.close()


@pytest.mark.parametrize("world,region_shift", [(2, None), (3, None), (3, 10)])
def test_distributed_build_on_one_gpu(gpu, world, region_shift, monkeypatch):
    """The distributed build (include/slacken_gpu.h, "Distributed build"; the shuffle of groupBy(idColumns).agg(udafLca),
    slacken/KeyValueIndex.scala:85-93) with one GPU playing every rank: rank r builds from the r-th block of genomes, reduces,
    hands over its cells grouped by owner; owner d receives the d-th group of every rank and inserts the runs. The shards'
    records must be exactly the oracle library's records, each on the rank slk_shard_of_records names, and a sharded
    classify over them must equal the oracle."""
    import torch
    from slacken_b200 import LibraryBuilder
    if region_shift is not None:   # 1 KB regions: the owner interleaves the runs region by region even on this small table
        monkeypatch.setenv("SLK_REGION_SHIFT", str(region_shift))
    rng, parents, ranks, names, genomes, taxa = make_world(53)
    olib = oracle_lib(oracle.params(), parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    params = IndexParams()
    dev = torch.device("cuda", gpu.device)
    sent, counts, dense = [], [], []
    for r in range(world):
        b = LibraryBuilder(gpu, tax, params)
        lo, hi = r * len(genomes) // world, (r + 1) * len(genomes) // world   # a strain and its relative meet on the owner
        gb, go = pack_sequences(genomes[lo:hi])
        b.add(gb, go, taxa[lo:hi])
        cnt = b.reduce(world)
        buf = b.cells_tensor()
        assert buf.numel() == sum(cnt) and buf.is_cuda
        dense.append(b.dense_taxa())
        sent.append(buf.clone()); counts.append(cnt)
        del buf
        b.close()
    assert all(d[0] == 0 for d in dense)
    owner = shard_of_records(params, id1, world)
    shards = []
    for d in range(world):
        runs = [sent[r][sum(counts[r][:d]):sum(counts[r][:d + 1])] for r in range(world)]
        recv = torch.cat(runs).contiguous()
        torch.cuda.synchronize()
        index = KeyValueIndex.from_cell_runs(gpu, tax, params, world, recv.data_ptr(), [c[d] for c in counts],
                                             np.concatenate(dense), [len(x) for x in dense])
        sid, stx = index.records()
        assert np.array_equal(sid, id1[owner == d]) and np.array_equal(stx, tx[owner == d])
        shards.append(index)
    union = np.unique(tx)
    ops = [GpuSplitOps(s, union) for s in shards]
    reads = simulate_reads(rng, genomes, 800, (30, 260), n_rate=0.1)
    rb, ro = pack_sequences(reads)
    q = ops[0]   # rank 0 asks, every shard answers
    d_b, d_o = q.upload(rb), q.upload(ro.view(np.int64))
    span_off, spans, n_spans = q.scan_spans(d_b, d_o, None, None, len(reads))
    keys, idx, kc = q.route(spans, n_spans, world)
    starts = np.concatenate([[0], np.cumsum(kc)])
    answers = torch.cat([ops[r].probe(keys[starts[r]:starts[r + 1]].contiguous()) for r in range(world)])
    got = q.resolve(spans, span_off, n_spans, len(reads), False, idx, answers, 0.1, 2, True)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), confidence=0.1)
    assert_batch_equal(res, per, got, 35)
    for o in ops:
        o.close()
    for s in shards:
        s.close()
    tax.close()
