"""world_size-2 gloo test of the sharded path's host logic: row shards -> local count tensors ->
the single exchange step (engine.Dist all-reduce) -> statistics identical to the unsharded run."""
import os
import socket

import numpy as np
import pytest

import parity_utils  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_counts(cols, doms, pairs):
    hist = [np.bincount(c.astype(np.int64) + 1, minlength=d + 1) for c, d in zip(cols, doms)]
    tabs = []
    for x, y in pairs:
        idx = (cols[x].astype(np.int64) + 1) * (doms[y] + 1) + cols[y] + 1
        tabs.append(np.bincount(idx, minlength=(doms[x] + 1) * (doms[y] + 1)))
    return hist, tabs


def _fd_tables(key, b, space):
    lo = np.full(space, 2 ** 31 - 1, dtype=np.int32)
    hi = np.full(space, -2 ** 31, dtype=np.int32)
    np.minimum.at(lo, key, (b + 1).astype(np.int32))
    np.maximum.at(hi, key, (b + 1).astype(np.int32))
    return lo, hi


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from repair import stats_host as SH
    from repair import synth
    from repair.engine import Dist
    from repair.table import EncodedTable
    n, k = 40000, 8
    spec = synth.SynthSpec.c4(n, k, seed=2)
    full = synth.generate_numpy(spec)
    table = EncodedTable.from_codes("tid", synth.column_names(k), full, spec.dom)
    shard = table.shard(rank, world)
    assert shard.n_rows_global == n and shard.row_offset == (n * rank) // world
    # the generator is shardable: this rank's rows straight from (seed, row, col)
    mine = synth.generate_numpy(spec, shard.row_offset, shard.row_offset + shard.n_rows)
    cols = [c.codes for c in shard.columns]
    assert all(np.array_equal(a, b) for a, b in zip(mine, cols))
    names, doms = table.names, spec.dom
    pairs = [(4, 5), (7, 6), (0, 1)]
    dist = Dist()
    hist, tabs = _local_counts(cols, doms, pairs)
    flat = torch.from_numpy(np.concatenate(hist + tabs).astype(np.int64))
    # FD key tables (MIN / MAX) and presence bits (OR) travel in the SAME exchange as the counts
    lo, hi = _fd_tables(cols[5].astype(np.int64) + 1, cols[4], doms[5] + 1)
    tlo, thi = torch.from_numpy(lo), torch.from_numpy(hi)
    idx = (cols[0].astype(np.int64) + 1) * (doms[1] + 1) + cols[1] + 1
    local_bits = np.zeros(((doms[0] + 1) * (doms[1] + 1) + 31) // 32, dtype=np.uint32)
    np.bitwise_or.at(local_bits, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
    tbits = torch.from_numpy(local_bits.view(np.int32).copy())
    # ... and so does the one-element "did the presence sample cover every shard" flag (MIN)
    cover = torch.tensor([1 if rank == 0 else 0], dtype=torch.int64)
    dist.exchange(None, [(flat, "sum"), (tlo, "min"), (thi, "max"), (tbits, "or"), (cover, "min")])   # THE exchange step
    assert dist.n_exchanges == 1 and int(cover.item()) == 0
    g = flat.numpy()
    ghist, off = {}, 0
    for nm, d in zip(names, doms):
        ghist[nm] = g[off:off + d + 1]
        off += d + 1
    gtabs = {}
    for (x, y) in pairs:
        e = (doms[x] + 1) * (doms[y] + 1)
        gtabs[(names[x], names[y])] = g[off:off + e].reshape(doms[x] + 1, doms[y] + 1)
        off += e
    ndv = {nm: d for nm, d in zip(names, doms)}
    named_pairs = [(names[x], names[y]) for x, y in pairs]
    got = SH.pairwise_entropies(n, ghist, gtabs, named_pairs, ndv)
    fhist, ftabs = _local_counts(full, doms, pairs)
    want = SH.pairwise_entropies(n, dict(zip(names, fhist)),
                                 {p: t.reshape(doms[x] + 1, doms[y] + 1) for p, t, (x, y) in
                                  zip(named_pairs, ftabs, pairs)}, named_pairs, ndv)
    assert got == want                                   # integer sums are order independent -> bit identical
    # FD key tables: MIN / MAX segments of the exchange
    flo, fhi = _fd_tables(full[5].astype(np.int64) + 1, full[4], doms[5] + 1)
    assert np.array_equal(tlo.numpy(), flo) and np.array_equal(thi.numpy(), fhi)
    viol = (flo != fhi)
    assert viol[-3:].all() and not viol[1:-3].any()      # exactly the three dirty determinant values
    fidx = (full[0].astype(np.int64) + 1) * (doms[1] + 1) + full[1] + 1
    fbits = np.zeros_like(local_bits)
    np.bitwise_or.at(fbits, fidx >> 5, np.uint32(1) << (fidx & 31).astype(np.uint32))
    assert np.array_equal(tbits.numpy().view(np.uint32), fbits)       # OR segment
    # every rank ingested its own rows with its own dictionaries: unify() makes the codes global
    import pandas as pd
    lo_r, hi_r = shard.row_offset, shard.row_offset + shard.n_rows
    strs = {nm: np.array(["v%03d" % v if v >= 0 else None for v in full[i][lo_r:hi_r]], dtype=object)
            for i, nm in enumerate(names[:3])}
    local = EncodedTable.from_pandas(pd.DataFrame({"tid": np.arange(lo_r, hi_r), **strs}), "tid")
    whole = EncodedTable.from_pandas(pd.DataFrame(
        {"tid": np.arange(n), **{nm: np.array(["v%03d" % v if v >= 0 else None for v in full[i]], dtype=object)
                                 for i, nm in enumerate(names[:3])}}), "tid")
    uni = local.unify(dist)
    assert uni.row_offset == lo_r and uni.n_rows_global == n
    for cu, cw in zip(uni.columns, whole.columns):
        assert list(cu.dictionary) == list(cw.dictionary) and np.array_equal(cu.codes, cw.codes[lo_r:hi_r])
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    td.destroy_process_group()


def test_two_rank_count_allreduce(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
