"""Host-side mirror of the reference's solid_geometry.py interface (magni, norm, plane, line, obstacle).

`obstacle.collis_det(vert_traj, horizon)` keeps the reference's contract (solid_geometry.py:104-168: score of
the first crossing of the gate plane by one rotor-tip polyline, `.co` flag) and evaluates it on the GPU through
`se3mpc_collis_det_host`; the fused path (4 tips + path + reward from the state trajectory) is `ops.reward`.
"""
from __future__ import annotations

import numpy as np


def magni(vector):
    v = np.asarray(vector, dtype=float)
    return np.sqrt(np.dot(v, v))


def norm(vector):
    v = np.asarray(vector, dtype=float)
    return v / magni(v)


class plane:
    """solid_geometry.py:16-47: a plane through three points; `normal` = unit(vec2 x vec1).  These small per-object helpers are kept
    for callers that poke at `obstacle.plane1.nor_vec()` etc.; the batched scoring never goes through them (the device code has its own
    gate set-up, se3mpc_core.cuh gate_setup / crossing_score)."""

    def __init__(self, point1, point2, point3):
        self.point1, self.point2, self.point3 = (np.asarray(q, float) for q in (point1, point2, point3))
        self.vec1, self.vec2 = self.point2 - self.point1, self.point3 - self.point1
        self.normal = norm(np.cross(self.vec2, self.vec1))

    def nor_vec(self):
        return self.normal

    def n1(self):
        return norm(np.cross(self.vec1, self.normal))

    def n2(self):
        return norm(np.cross(self.normal, self.vec2))

    def n3(self):
        self.vec3 = self.point3 - self.point2
        return norm(np.cross(self.normal, self.vec3))

    def interpoint(self, point1, point2):
        """Intersection of the plane with the line through point1 and point2 (solid_geometry.py:43-47)."""
        a, b = np.asarray(point1, float), np.asarray(point2, float)
        d = norm(a - b)
        return a - (np.dot(self.normal, a - self.point1) / np.dot(d, self.normal)) * d


class line:
    """solid_geometry.py:50-78: the segment point1-point2; `vertical` = distance to the infinite line, `distance` = to the segment."""

    def __init__(self, point1, point2):
        self.point1, self.point2 = np.asarray(point1, float), np.asarray(point2, float)
        self.dir = norm(self.point1 - self.point2)

    def vertical(self, point):
        return magni(np.cross(np.asarray(point, float) - self.point1, self.dir))

    def distance(self, point):
        q = np.asarray(point, float)
        a, b, c, d = self.vertical(q), magni(q - self.point1), magni(q - self.point2), magni(self.point1 - self.point2)
        far, near = (b, c) if b > c else (c, b)          # the foot of the perpendicular lies outside the segment iff far^2 - d^2 > a^2
        return near if (far ** 2 - d ** 2) > a ** 2 else a


class obstacle:
    def __init__(self, point1, point2, point3, point4):
        self.point1, self.point2 = np.asarray(point1, float), np.asarray(point2, float)
        self.point3, self.point4 = np.asarray(point3, float), np.asarray(point4, float)
        self.centroid = (self.point1 + self.point2 + self.point3 + self.point4) / 4
        pts = (self.point1, self.point2, self.point3, self.point4)
        for i in range(4):                                # solid_geometry.py:93-102
            setattr(self, f"plane{i + 1}", plane(self.centroid, pts[i], pts[(i + 1) % 4]))
            setattr(self, f"line{i + 1}", line(pts[i], pts[(i + 1) % 4]))
        self.gate12 = np.concatenate(pts)
        self.co = 0

    def collis_det(self, vert_traj, horizon):
        from . import ops
        V = np.asarray(vert_traj, dtype=float).reshape(-1, 3)
        score, co = ops.collis_det_host(V[None, : horizon + 1], self.gate12[None], horizon)
        self.co = int(co[0])
        return float(score[0])
