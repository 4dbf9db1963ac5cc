"""Host-side mirror of the reference's solid_geometry.py interface (magni, norm, obstacle).

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


class obstacle:
    def __init__(self, point1, point2, point3, point4):
        self.point1, self.point2 = np.asarray(point1, float), np.asarray(point2, float)
        self.point3, self.point4 = np.asarray(point3, float), np.asarray(point4, float)
        self.centroid = (self.point1 + self.point2 + self.point3 + self.point4) / 4
        self.gate12 = np.concatenate([self.point1, self.point2, self.point3, self.point4])
        self.co = 0

    def collis_det(self, vert_traj, horizon):
        from . import ops
        V = np.asarray(vert_traj, dtype=float).reshape(-1, 3)
        score, co = ops.collis_det_host(V[None, : horizon + 1], self.gate12[None], horizon)
        self.co = int(co[0])
        return float(score[0])
