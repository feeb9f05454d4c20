"""CPU: the inference driver against golden output of the REFERENCE driver
(tests/golden/make_inference_golden.py ran core/unopose/engine/oneref_inference_utils_v1.py on the same fake
model / loader), single process and sharded across 2 gloo ranks."""
import json
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from inference_fixtures import FakePoseModel, make_loader  # noqa: E402
from unopose_b200 import inference as INF  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "inference_v1.json")))


def _run(tmp_path, with_tem, bs=16, group=None, shard=None, first=None):
    model = FakePoseModel()
    path = os.path.join(str(tmp_path), "result.csv")
    INF.time.perf_counter = lambda: 0.0          # the `time` column becomes the deterministic seg_time
    try:
        loader = make_loader(0, 3, with_tem)
        for d in (loader[first:] if first is not None else []):                 # `first` > 0: this rank's loader yields other images from there on
            d["img_id"] = d["img_id"] + 1000
        INF.inference_and_save_oneref_v1(model, loader, path, instance_batch_size=bs, group=group, shard_instances=shard)
    finally:
        import time as _t

        INF.time.perf_counter = _t.perf_counter
    return model, path


def test_matches_reference_driver_byte_for_byte(tmp_path):
    for name, tem in (("with_tem_pose", True), ("plain", False)):
        model, path = _run(tmp_path, tem)
        assert open(path).read() == GOLD[name]["csv"], name
        assert json.loads(open(path.replace(".csv", ".json")).read()) == GOLD[name]["json"], name
        assert model.calls == [5, 16, 16, 5, 16]     # chunks of the 5 / 37 / 16-instance images


def test_chunk_size_does_not_change_results(tmp_path):
    _, p1 = _run(tmp_path, True, bs=16)
    a = open(p1).read()
    _, p2 = _run(tmp_path, True, bs=7)
    assert open(p2).read() == a
    assert INF.instance_chunks(3, 40, 16) == [(3, 19), (19, 35), (35, 40)]
    assert INF.instance_chunks(5, 5, 16) == []


def test_line_format():
    line = INF.format_result_line(48, 1, 14, 0.25, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0], [1.5, 2.0, 3.0], 0.125)
    assert line == "48,1,14,0.25,1.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 1.0,1.5 2.0 3.0,0.125\n"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class P:
            def __str__(self):
                return os.path.join(tmp, "rank%d" % rank)
        os.makedirs(str(P()), exist_ok=True)
        model, path = _run(P(), True, shard=True)
        # instances of every image were split 3+2 / 19+18 / 8+8 between the two ranks
        res = (model.calls, open(path).read() if rank == 0 else os.path.exists(path))
        # default (no group, no opt-in): every rank behaves like the reference driver, no collective, all instances
        model2, path2 = _run(P(), True)
        res += (model2.calls, open(path2).read())
        # sharding with DIFFERENT images per rank (what the reference's InferenceSampler would hand out) must raise
        try:
            _run(P(), True, shard=True, first=3 - 2 * rank)
            res += (False,)
        except RuntimeError as ex:
            res += ("same image on every rank" in str(ex),)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_world2_gloo_instance_partition(tmp_path):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), out), nprocs=world, join=True)
    calls0, csv0, dcalls0, dcsv0, raised0 = out[0]
    calls1, wrote1, dcalls1, dcsv1, raised1 = out[1]
    assert csv0 == GOLD["with_tem_pose"]["csv"]          # gathered result identical to one process / the reference
    assert wrote1 is False                                # only rank 0 writes
    assert calls0 == [3, 16, 3, 8] and calls1 == [2, 16, 2, 8]
    assert dcalls0 == dcalls1 == [5, 16, 16, 5, 16] and dcsv0 == dcsv1 == GOLD["with_tem_pose"]["csv"]
    assert raised0 is True and raised1 is True
