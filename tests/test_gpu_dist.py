"""Row-sharded run == one-GPU run, on the device, through the public API
(``RepairModel.setDistributed``): error cells, trained models (same global training sample) and
repairs.  Two ranks: over NCCL when the box has two GPUs, else both ranks share cuda:0 and exchange
through gloo -- the sharded code path (shard-local kernels, packed exchange + dr_combine_counts,
global sample, dictionary unification) is the same."""
import os
import socket

import numpy as np
import pytest

import parity_utils as PU

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

N_ROWS, N_COLS = 200_000, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _specs():
    from repair import synth
    return [{"type": "null"}, {"type": "constraint", "constraints": synth.fd_constraints(N_COLS)},
            # a general two-tuple constraint: its projection bits travel in the same exchange
            {"type": "constraint", "constraints": "t1&t2&EQ(t1.c01,t2.c01)&IQ(t1.c02,t2.c02)&IQ(t1.c00,t2.c00)",
             "targets": ["c00"]}]


OPTS = {"error.pairwise_freq_ratio_threshold": 1.0, "model.lgb.n_estimators": 8, "model.max_training_row_num": 3000,
        "model.hp.max_evals": 1}


def _run(table, dist, device_index, mode):
    from repair import RepairModel
    rm = RepairModel().setEncodedInput(table).setErrorDetectors(PU.make_detectors(_specs()))
    for k, v in OPTS.items():
        rm.option(k, str(v))
    if dist:
        rm.setDistributed(True, device_index)
    out = rm.run(detect_errors_only=(mode == "detect"))
    return rm, PU.frame_tuples(out, "tid")


def _worker(rank, world, port, backend, out_dir):
    import torch.distributed as td
    from repair import synth
    from repair.table import EncodedTable
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    td.init_process_group(backend, rank=rank, world_size=world)
    spec = synth.SynthSpec.c4(N_ROWS, N_COLS, seed=3)
    names = synth.column_names(N_COLS)
    lo, hi = (N_ROWS * rank) // world, (N_ROWS * (rank + 1)) // world
    mine = synth.generate_numpy(spec, lo, hi)
    shard = EncodedTable.from_codes("tid", names, mine, spec.dom, row_ids=np.arange(lo, hi, dtype=np.int64))
    result = {}
    for mode in ("detect", "repair"):
        rm, got = _run(shard, True, dev, mode)
        gathered = [None] * world
        td.all_gather_object(gathered, got)
        result[mode] = sorted((t for part in gathered for t in part), key=lambda t: (int(t[0]), t[1]))
        assert rm.last_run["gpu_launches"] > 0
    # the same shard as a pyarrow.Table whose dictionaries are rank-local (first-seen order, only the values
    # the shard holds): device-side ingest + dictionary unification on the device must give the same repairs
    import pyarrow as pa
    import pyarrow.compute as pc
    from repair import RepairModel
    cols = {"tid": pa.array(np.arange(lo, hi, dtype=np.int64))}
    for nm, c in zip(names, mine):
        strs = pa.array([None if v < 0 else "v%03d" % v for v in c.tolist()], type=pa.string())
        cols[nm] = pc.dictionary_encode(strs)
    rma = RepairModel().setArrowInput(pa.table(cols)).setRowId("tid").setErrorDetectors(PU.make_detectors(_specs()))
    for k, v in OPTS.items():
        rma.option(k, str(v))
    frame = rma.setDistributed(True, dev).run()
    assert isinstance(frame, pa.Table)
    assert PU.frame_tuples(frame.to_pandas(), "tid") == got, "Arrow ingest of the shard differs from the encoded input"
    if rank == 0:
        full = EncodedTable.from_codes("tid", names, synth.generate_numpy(spec), spec.dom)
        for mode in ("detect", "repair"):
            _, want = _run(full, False, dev, mode)
            want = sorted(want, key=lambda t: (int(t[0]), t[1]))
            assert len(want) > 1000
            assert result[mode] == want, mode
        # every rank emitted rows of its own shard only
    assert all(lo <= int(t[0]) < hi for t in got)
    td.barrier()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    td.destroy_process_group()


def test_two_rank_sharded_run_equals_one_gpu_run(tmp_path):
    import torch.multiprocessing as mp
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    mp.spawn(_worker, args=(2, _free_port(), backend, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
