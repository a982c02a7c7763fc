"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the supervision geometry of the AVDN training rollout
(SURVEY.md §8f N3): ``compute_iou`` (src/xview_et/agent.py:46-78) and ``teacher_action`` with student feedback
(src/xview_et/agent.py:386-507).  The reference evaluates both with shapely / GEOS, which is absent from this
image: **parity with shapely is unpinned**.  The geometry itself is pinned independently in
tests/test_teacher_oracle.py: the convex intersection and hull areas against OpenCV (``cv2.intersectConvexConvex``,
``cv2.convexHull`` + ``cv2.contourArea``) and the segment/polygon exit point against a bisection on a
point-in-polygon test.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import
this module.

What the shapely calls compute, for the convex quadrilaterals the simulator produces:
  * ``Polygon(a).convex_hull`` / ``.intersection(...).area``   -> area of the intersection of two convex quads
  * ``MultiPoint(a + b).convex_hull.area``                     -> area of the convex hull of the 8 corners
    (the reference's "IoU" divides by THIS, not by the union: agent.py:64-69)
  * ``Polygon(corners).intersection(LineString([cur, goal])).coords`` with ``cur`` the quad's centre -> the part of
    the segment inside the quad: ``[cur, goal]`` if the goal is inside, else ``[cur, exit point]``.
"""
import numpy as np


def _area(poly):
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _ccw(poly):
    return poly if _area(poly) >= 0 else poly[::-1]


def convex_hull(points):
    """Andrew's monotone chain; returns the hull counter-clockwise."""
    pts = sorted(map(tuple, np.asarray(points, dtype=np.float64)))
    if len(pts) <= 2:
        return np.array(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


def clip_convex(subject, clip):
    """Sutherland-Hodgman: intersection polygon of two convex polygons (both made CCW first)."""
    out = [tuple(p) for p in _ccw(np.asarray(subject, dtype=np.float64))]
    cl = _ccw(np.asarray(clip, dtype=np.float64))
    for i in range(len(cl)):
        a, b = cl[i], cl[(i + 1) % len(cl)]
        inp, out = out, []
        if not inp:
            break

        def side(p):
            return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        for j in range(len(inp)):
            p, q = inp[j], inp[(j + 1) % len(inp)]
            sp, sq = side(p), side(q)
            if sp >= 0:
                out.append(p)
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return np.array(out) if out else np.zeros((0, 2))


def compute_iou(a, b):
    """agent.py:46-78: intersection area / area of the convex hull of all eight corners (0 if disjoint)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    inter = clip_convex(convex_hull(a), convex_hull(b))
    if len(inter) < 3:
        return 0.0
    ia = abs(_area(inter))
    ua = abs(_area(convex_hull(np.concatenate((a, b)))))
    return 0.0 if ua == 0 else ia / ua


def inside_convex(poly, p):
    poly = _ccw(np.asarray(poly, dtype=np.float64))
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        if (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) < 0:
            return False
    return True


def segment_exit(poly, cur, goal):
    """Last point of the segment cur -> goal that lies inside the convex polygon (cur is inside): the goal itself
    if it is inside, else the boundary crossing."""
    poly = _ccw(np.asarray(poly, dtype=np.float64))
    cur, goal = np.asarray(cur, dtype=np.float64), np.asarray(goal, dtype=np.float64)
    t_exit = 1.0
    d = goal - cur
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        e = b - a
        s0 = e[0] * (cur[1] - a[1]) - e[1] * (cur[0] - a[0])        # >= 0: cur on the inner side
        s1 = e[0] * (goal[1] - a[1]) - e[1] * (goal[0] - a[0])
        if s1 < 0 <= s0:
            t_exit = min(t_exit, s0 / (s0 - s1))
    return cur + t_exit * d


def clip_segment(poly, p, q):
    """The part of the segment p -> q inside the convex polygon: (a, b) or None (Cyrus-Beck)."""
    poly = _ccw(np.asarray(poly, dtype=np.float64))
    t0, t1 = 0.0, 1.0
    d = q - p
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        e = b - a
        s0 = e[0] * (p[1] - a[1]) - e[1] * (p[0] - a[0])
        ds = e[0] * d[1] - e[1] * d[0]                    # d(side)/dt
        if ds == 0:
            if s0 < 0:
                return None
            continue
        t = -s0 / ds
        if ds > 0:
            t0 = max(t0, t)                               # entering
        else:
            t1 = min(t1, t)                               # leaving
    if t0 > t1:
        return None
    return p + t0 * d, p + t1 * d


def path_candidates(poly, line):
    """Coordinates of ``Polygon(poly).intersection(LineString(line))`` (agent.py:441-449): for a convex polygon the
    union of the clipped segments -- their end points are the boundary crossings and the polyline vertices inside."""
    out = []
    for p, q in zip(line[:-1], line[1:]):
        c = clip_segment(poly, np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64))
        if c is not None:
            out += [c[0], c[1]]
    return out


def teacher_action(corners, gt_path_corners, ended, feedback="student"):
    """agent.py:386-507 for ONE sample.  corners [4,2] (lat, lng), gt_path_corners [n,4,2].
    ``feedback == 'student'``: the target is where the segment view centre -> goal centre leaves the view (the goal
    itself if it is inside); ``'teacher'``: the point of (ground-truth path ∩ view) closest to the goal, falling back
    to the student rule when the path misses the view (agent.py:451-456).
    Returns (next_pos_ratio float32 [2], altitude float, progress float32)."""
    corners = np.asarray(corners, dtype=np.float64)
    gt = np.asarray(gt_path_corners, dtype=np.float64)
    cur = np.mean(corners, axis=0)
    progress = np.float32(compute_iou(corners, gt[-1]))
    min_dis, closest = 1000.0, 0
    for j in range(len(gt) - 1, -1, -1):
        dis = np.linalg.norm(np.mean(gt[j], axis=0) - cur)
        if dis + 0.00001 < min_dis:
            min_dis, closest = dis, j
    altitude = float((np.linalg.norm(gt[closest][0] - gt[closest][1]) * 11.13 * 1e4 - 40) / (400 - 40))
    if ended or progress > 0.5:
        return np.array([0, 0], dtype=np.float32), altitude, progress
    goal = np.mean(gt[-1], axis=0)
    x = None
    if feedback == "teacher":
        best = 1.0                                         # min_distance = 1 (agent.py:461)
        for c in path_candidates(corners, [np.mean(g, axis=0) for g in gt]):
            dist = np.linalg.norm(c - goal)
            if dist < best:
                best, x = dist, c
    if x is None:
        x = segment_exit(corners, cur, goal)              # the coords of the intersection closest to the goal
    net_next = 1e5 * (x - cur)
    net_y = np.round(1e5 * ((corners[0] + corners[1]) / 2 - cur)).astype(np.int64)
    net_x = np.round(1e5 * ((corners[1] + corners[2]) / 2 - cur)).astype(np.int64)
    A = np.array([[net_x[0], net_y[0]], [net_x[1], net_y[1]]], dtype=np.float64)
    r = np.linalg.solve(A, net_next.reshape(2, 1)).reshape(2)
    m = max(abs(r[0]), abs(r[1]), 1)
    return np.array([r[0] / m, r[1] / m], dtype=np.float32), altitude, progress
