"""Small numpy SE3 helpers (tangent order [upsilon; omega] like Sophus) for the synthetic generators and tests."""
from __future__ import annotations

import numpy as np


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)


def so3_exp(w):
    th = float(np.linalg.norm(w))
    K = hat(w)
    if th < 1e-10:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def so3_log(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-10:
        return np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def se3_exp(v):
    """v = [upsilon; omega] -> 3x4 [R|t]."""
    u, w = np.asarray(v[:3], np.float64), np.asarray(v[3:], np.float64)
    th = float(np.linalg.norm(w))
    K = hat(w)
    R = so3_exp(w)
    if th < 1e-10:
        V = np.eye(3) + 0.5 * K
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * (K @ K)
    T = np.zeros((3, 4))
    T[:, :3] = R
    T[:, 3] = V @ u
    return T


def se3_log(T):
    R, t = T[:, :3], T[:, 3]
    w = so3_log(R)
    th = float(np.linalg.norm(w))
    K = hat(w)
    if th < 1e-10:
        Vi = np.eye(3) - 0.5 * K
    else:
        Vi = np.eye(3) - 0.5 * K + (1 - th / (2 * np.tan(th / 2))) / th**2 * (K @ K)
    return np.concatenate([Vi @ t, w])


def mul(A, B):
    """compose two 3x4 transforms: A * B"""
    T = np.zeros((3, 4))
    T[:, :3] = A[:, :3] @ B[:, :3]
    T[:, 3] = A[:, :3] @ B[:, 3] + A[:, 3]
    return T


def inv(A):
    T = np.zeros((3, 4))
    T[:, :3] = A[:, :3].T
    T[:, 3] = -A[:, :3].T @ A[:, 3]
    return T
