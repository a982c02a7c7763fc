"""Pins ``ANDHNavBatch._eval_item`` / ``eval_metrics`` against the REFERENCE's own code
(src/env.py:335-475), run in the build container.  shapely is absent here: ``Point`` / ``Polygon.contains`` are
stubbed with an INDEPENDENT third-party point-in-polygon test (``cv2.pointPolygonTest`` on coordinates centred and
scaled to float32-friendly units); everything else -- lengths, goal progress, success, SPL, the grouping and the
averages -- is the reference's code.  Writes tests/golden/eval_golden.npz.

    python tests/golden/make_eval_golden.py
"""
import json
import os
import sys
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Point:
    def __init__(self, c):
        self.c = np.asarray(c, dtype=np.float64)


class Polygon:
    def __init__(self, pts):
        self.p = np.asarray(pts, dtype=np.float64)

    def contains(self, pt):
        c0 = self.p.mean(0)
        poly = ((self.p - c0) * 1e5).astype(np.float32).reshape(-1, 1, 2)
        q = (pt.c - c0) * 1e5
        return cv2.pointPolygonTest(poly, (float(q[0]), float(q[1])), False) > 0


geom = types.ModuleType("shapely.geometry")
geom.Point, geom.Polygon, geom.LineString, geom.MultiPoint = Point, Polygon, object, object
sh = types.ModuleType("shapely")
sh.geometry = geom
ops = types.ModuleType("shapely.ops")
ops.nearest_points = None
sys.modules.update({"shapely": sh, "shapely.geometry": geom, "shapely.ops": ops})
sys.path.insert(0, "/root/reference/src")
import env as ref_env  # noqa: E402


def quad(rng, c, side):
    th = rng.uniform(0, 2 * np.pi)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return c + (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * side / 2) @ R.T


def main():
    rng = np.random.default_rng(0)
    base = np.array([40.01, -74.99])
    preds = {}
    for i in range(40):
        n_gt, n_path = int(rng.integers(2, 6)), int(rng.integers(1, 7))
        c = base + rng.uniform(-0.004, 0.004, 2)
        gt = []
        for _ in range(n_gt):
            c = c + rng.uniform(-0.0015, 0.0015, 2)
            gt.append(quad(rng, c, rng.uniform(0.001, 0.003)))
        p = gt[0].mean(0)
        path = [(gt[0].copy(), 0)]
        for _ in range(n_path - 1):
            p = p + rng.uniform(-0.0015, 0.0015, 2)
            path.append((quad(rng, p, rng.uniform(0.001, 0.003)), int(rng.integers(0, 360))))
        if i % 3 == 0:                                   # ends on the goal: success candidates
            path.append((gt[-1] + rng.uniform(-2e-5, 2e-5, 2), 0))
        prog = [float(x) for x in rng.uniform(0, 1, size=len(path))]
        if i % 3 == 0:
            prog[-1] = float(rng.uniform(0.4, 1.0))
        preds[f"id{i}"] = dict(instr_id=f"id{i}", num_dia=int(rng.integers(1, 4)), path_corners=path,
                               gt_path_corners=gt, gt_progress=prog)
    e = ref_env.ANDHNavBatch.__new__(ref_env.ANDHNavBatch)
    avg, metrics = e.eval_metrics(preds)
    ha = {f"h{i}": dict(human_att_performance=[[float(a), float(b)] for a, b in rng.uniform(0, 1, size=(3, 2))],
                        nss=[float(x) for x in rng.uniform(0, 2, size=3)]) for i in range(5)}
    avg_h, _ = e.eval_metrics(ha, human_att_eval=True)
    blob = dict(preds={k: dict(instr_id=v["instr_id"], num_dia=v["num_dia"], gt_progress=v["gt_progress"],
                               path_corners=[[np.asarray(c).tolist(), int(d)] for c, d in v["path_corners"]],
                               gt_path_corners=[np.asarray(g).tolist() for g in v["gt_path_corners"]])
                       for k, v in preds.items()},
                avg={k: float(v) for k, v in avg.items()},
                per_item={k: [float(x) for x in metrics[k]] for k in ("trajectory_lengths", "gp", "oracle_gp", "success",
                                                                       "oracle_success", "spl", "gt_length", "iou")},
                ha=ha, avg_h={k: float(v) for k, v in avg_h.items()})
    out = os.path.join(ROOT, "tests", "golden", "eval_golden.json")
    with open(out, "w") as f:
        json.dump(blob, f)
    print("wrote", out, "sr", avg["sr"], "spl", avg["spl"], os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
