"""CPU: evaluation metrics and multi-GPU result merging -- ``ANDHNavBatch.eval_metrics`` against
the golden produced by the REFERENCE's own ``eval_metrics`` (tests/golden/make_eval_golden.py: src/env.py:335-475
with shapely's ``contains`` stubbed by cv2.pointPolygonTest), and the gather / merge helpers under gloo."""
import json
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avdn_b200 import parallel
from avdn_b200.env import ANDHNavBatch


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "eval_golden.json")) as f:
        g = json.load(f)
    preds = {}
    for k, v in g["preds"].items():
        preds[k] = dict(instr_id=v["instr_id"], num_dia=v["num_dia"], gt_progress=v["gt_progress"],
                        path_corners=[(np.array(c), d) for c, d in v["path_corners"]],
                        gt_path_corners=[np.array(x) for x in v["gt_path_corners"]])
    return g, preds


def test_eval_metrics_match_the_reference(golden_dir):
    g, preds = _golden(golden_dir)
    env = ANDHNavBatch.__new__(ANDHNavBatch)            # the metrics need no device
    avg, metrics = env.eval_metrics(preds)
    assert set(avg) == set(g["avg"])
    for k, v in g["avg"].items():
        assert abs(float(avg[k]) - v) <= 1e-9 * max(1.0, abs(v)), (k, avg[k], v)
    for k, v in g["per_item"].items():
        np.testing.assert_allclose(np.asarray(metrics[k], dtype=np.float64), np.asarray(v), rtol=1e-12, atol=1e-12, err_msg=k)
    assert 0 < avg["sr"] < 100                          # both success outcomes occur in the fixture
    avg_h, _ = env.eval_metrics(g["ha"], human_att_eval=True)
    for k, v in g["avg_h"].items():
        assert abs(float(avg_h[k]) - v) <= 1e-12, k


def test_contains_is_strict():
    q = np.array([[0, 0], [2, 0], [2, 2], [0, 2]], dtype=np.float64)
    c = ANDHNavBatch._contains
    assert c(q, (1, 1)) and c(q[::-1], (1, 1))
    assert not c(q, (2, 1)) and not c(q, (3, 1)) and not c(q, (0, 0))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = [dict(instr_id=f"r{rank}_{i}", path=np.full((2, 2), rank + i / 10)) for i in range(rank + 2)]
        merged = parallel.merge_dist_results(parallel.all_gather(mine))
        ok = [m["instr_id"] for m in merged] == ["r0_0", "r0_1", "r1_0", "r1_1", "r1_2"]
        ok = ok and float(merged[-1]["path"][0, 0]) == 1.2
        red = parallel.reduce_dict({"b": torch.tensor(float(rank + 1)), "a": torch.tensor(10.0 * rank)})
        ok = ok and float(red["a"]) == 5.0 and float(red["b"]) == 1.5
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_merge():
    assert parallel.all_gather({"x": 1}) == [{"x": 1}] and parallel.get_world_size() == 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
