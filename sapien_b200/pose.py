"""Minimal stand-in for ``sapien.Pose`` (reference: include/sapien/math/pose.h:24-47,
quat.h:23-82, conversion.h:38-51): float32 position + (w, x, y, z) quaternion, composition,
inverse and 4x4 matrix round trip.  Only what the stereo sensor calibration needs."""
from __future__ import annotations

import numpy as np


def _quat_mul(a, b):
    w, x, y, z = a
    qw, qx, qy, qz = b
    return np.array(
        [w * qw - x * qx - y * qy - z * qz, w * qx + qw * x + y * qz - qy * z,
         w * qy + qw * y + z * qx - qz * x, w * qz + qw * z + x * qy - qx * y], dtype=np.float32)


def _quat_rotate(q, v):
    w = q[0]
    u = q[1:]
    return (np.float32(2.0) * np.dot(u, v) * u + (w * w - np.dot(u, u)) * v
            + np.float32(2.0) * w * np.cross(u, v)).astype(np.float32)


def _mat_to_quat(m):
    m = np.asarray(m, dtype=np.float32)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q, dtype=np.float32)
    return q / np.linalg.norm(q)


class Pose:
    def __init__(self, p=None, q=None):
        if p is not None and np.asarray(p).shape == (4, 4):
            m = np.asarray(p, dtype=np.float32)
            self.q = _mat_to_quat(m[:3, :3])
            self.p = m[:3, 3].astype(np.float32).copy()
            return
        self.p = np.zeros(3, np.float32) if p is None else np.asarray(p, dtype=np.float32).reshape(3)
        self.q = np.array([1, 0, 0, 0], np.float32) if q is None else np.asarray(q, dtype=np.float32).reshape(4)

    def inv(self) -> "Pose":
        qc = self.q * np.array([1, -1, -1, -1], np.float32)
        return Pose(_quat_rotate(qc, -self.p), qc)

    def __mul__(self, other: "Pose") -> "Pose":
        return Pose(_quat_rotate(self.q, other.p) + self.p, _quat_mul(self.q, other.q))

    def to_transformation_matrix(self) -> np.ndarray:
        w, x, y, z = (self.q / np.linalg.norm(self.q)).astype(np.float32)
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = np.array(
            [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
             [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
             [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float32)
        m[:3, 3] = self.p
        return m

    def __repr__(self):
        return f"Pose({self.p.tolist()}, {self.q.tolist()})"
