"""CPU: the supervision-geometry restatement (oracle/teacher_oracle.py; shapely is absent, see its header) pinned
independently -- intersection / hull areas against OpenCV, the exit point against bisection."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import teacher_oracle as to


def _quad(rng, c=None, size=None):
    c = rng.uniform(-1, 1, 2) if c is None else c
    w, h = (rng.uniform(0.3, 1.5, 2) if size is None else size)
    th = rng.uniform(0, 2 * np.pi)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return c + (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * np.array([w, h]) / 2) @ R.T


def test_iou_matches_opencv():
    rng = np.random.default_rng(0)
    n_pos = 0
    for _ in range(300):
        a, b = _quad(rng), _quad(rng, c=rng.uniform(-1.5, 1.5, 2))
        ia, _ = cv2.intersectConvexConvex(a.astype(np.float32), b.astype(np.float32))
        hull = cv2.convexHull(np.concatenate((a, b)).astype(np.float32))
        ref = ia / cv2.contourArea(hull) if ia > 0 else 0.0
        ours = to.compute_iou(a, b)
        assert abs(ours - ref) < 2e-5, (ours, ref)
        n_pos += ours > 0
    assert 50 < n_pos < 300
    q = _quad(rng)
    assert abs(to.compute_iou(q, q) - 1.0) < 1e-12
    assert to.compute_iou(q, q + 10.0) == 0.0
    # axis-aligned closed form: unit squares shifted by (0.5, 0): inter 0.5, hull 1.5
    s = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float64)
    assert abs(to.compute_iou(s, s + np.array([0.5, 0.0])) - 0.5 / 1.5) < 1e-12


def test_segment_exit_matches_bisection():
    rng = np.random.default_rng(1)
    for _ in range(200):
        q = _quad(rng)
        cur = q.mean(0)
        goal = cur + rng.uniform(-2, 2, 2)
        x = to.segment_exit(q, cur, goal)
        if to.inside_convex(q, goal):
            assert np.allclose(x, goal)
            continue
        lo, hi = 0.0, 1.0
        for _ in range(80):
            mid = (lo + hi) / 2
            if to.inside_convex(q, cur + mid * (goal - cur)):
                lo = mid
            else:
                hi = mid
        assert np.allclose(x, cur + lo * (goal - cur), atol=1e-12)


def test_teacher_action_semantics():
    rng = np.random.default_rng(2)
    base = np.array([40.01, -74.99])
    q = base + _quad(rng, c=np.zeros(2), size=(0.004, 0.004))
    far = base + np.array([0.01, 0.004]) + _quad(rng, c=np.zeros(2), size=(0.003, 0.003))
    r, alt, prog = to.teacher_action(q, [q, far], ended=False)
    assert prog == 0 and r.dtype == np.float32 and np.max(np.abs(r)) <= 1.0 + 1e-6
    assert abs(np.max(np.abs(r)) - 1.0) < 0.02            # goal outside the view: the target sits on the view's edge
    r2, _, prog2 = to.teacher_action(q, [far, q], ended=False)
    assert prog2 > 0.99 and np.all(r2 == 0)               # already on the goal: progress > 0.5 -> zero action
    r3, _, _ = to.teacher_action(q, [q, far], ended=True)
    assert np.all(r3 == 0)
    assert abs(alt - (np.linalg.norm(q[0] - q[1]) * 11.13e4 - 40) / 360) < 1e-9


def test_clip_segment_and_teacher_feedback():
    rng = np.random.default_rng(4)
    for _ in range(200):
        q = _quad(rng)
        p0, p1 = rng.uniform(-2, 2, 2), rng.uniform(-2, 2, 2)
        c = to.clip_segment(q, p0, p1)
        ts = np.linspace(0, 1, 2001)
        inside = np.array([to.inside_convex(q, p0 + t * (p1 - p0)) for t in ts])
        if c is None:
            assert inside.sum() <= 1
            continue
        a, b = c
        ta, tb = np.dot(a - p0, p1 - p0) / np.dot(p1 - p0, p1 - p0), np.dot(b - p0, p1 - p0) / np.dot(p1 - p0, p1 - p0)
        if inside.any():
            assert abs(ts[inside].min() - ta) < 1e-3 and abs(ts[inside].max() - tb) < 1e-3
    # teacher feedback: the path crosses the view; the target is the crossing nearest to the goal
    base = np.array([40.01, -74.99])
    view = base + _quad(rng, c=np.zeros(2), size=(0.004, 0.004))
    g0 = view - np.array([0.006, 0.0])
    g1 = view + np.array([0.006, 0.001])
    rt, _, _ = to.teacher_action(view, [g0, g1], ended=False, feedback="teacher")
    rs, _, _ = to.teacher_action(view, [g0, g1], ended=False, feedback="student")
    assert np.max(np.abs(rt)) <= 1 + 1e-6 and np.max(np.abs(rs)) <= 1 + 1e-6
    # a path that misses the view falls back to the student rule
    far0, far1 = view + np.array([0.02, 0.02]), view + np.array([0.03, 0.02])
    ft, _, _ = to.teacher_action(view, [far0, far1], ended=False, feedback="teacher")
    fs, _, _ = to.teacher_action(view, [far0, far1], ended=False, feedback="student")
    assert np.array_equal(ft, fs)
