"""Deterministic synthetic clouds for the parity tests and the benchmark.

SURVEY.md §8(d): counter-based RNG (splitmix64 -> 24-bit mantissa uniforms) so any
language reproduces the same bits; scene S(P, seed) in a KITTI/demo-like slab;
source = N base points + noise; target = M points with 80 % index overlap, moved by
T_gt^-1 with T_gt = Rot_y(2 deg) * Trans(0.05, 0.02, 0.50).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, stream: int, n: int) -> np.ndarray:
    """n float32 uniforms in [0,1) from (seed, stream, counter)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed], dtype=np.uint64) * np.uint64(0x100000001B3)
                           + np.uint64(stream))[0]
        ctr = np.arange(n, dtype=np.uint64) + base
    bits = _splitmix64(ctr) >> np.uint64(40)  # top 24 bits
    return (bits.astype(np.float64) / float(1 << 24)).astype(np.float32)


def normal(seed: int, stream: int, n: int) -> np.ndarray:
    u1 = uniform01(seed, 2 * stream + 1000, n).astype(np.float64)
    u2 = uniform01(seed, 2 * stream + 1001, n).astype(np.float64)
    u1 = np.maximum(u1, 1.0 / (1 << 24))
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)


def gt_transform() -> np.ndarray:
    """T_gt = Rot_y(2 deg) * Trans(0.05, 0.02, 0.50), float64 4x4."""
    a = np.deg2rad(2.0)
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Tr = np.eye(4)
    Tr[:3, 3] = [0.05, 0.02, 0.50]
    Rm = np.eye(4)
    Rm[:3, :3] = R
    return Rm @ Tr


def make_pair(P: int, N: int, M: int, seed: int, F: int = 0, C: int = 0, overlap: float = 0.8,
              with_geotype: bool = False):
    """Returns dict(source=..., target=..., T_gt=...), each cloud a dict of float32 arrays."""
    n_ov = min(int(round(overlap * M)), N)
    assert P >= N + (M - n_ov), "base scene too small for the requested overlap"
    bx = uniform01(seed, 1, P) * 20.0 - 10.0
    by = uniform01(seed, 2, P) * 4.0 - 2.0
    bz = uniform01(seed, 3, P) * 28.0 + 2.0
    base = np.stack([bx, by, bz], axis=1).astype(np.float32)
    feat = None
    if F > 0:
        cols = [uniform01(seed, 10 + k, P) for k in range(min(F, 3))]
        for k in range(3, F):
            cols.append(np.clip(0.5 + 0.1 * normal(seed, 20 + k, P), 0.0, 1.0).astype(np.float32))
        feat = np.stack(cols, axis=1).astype(np.float32)
    lab = None
    if C > 0:
        cls = np.minimum((uniform01(seed, 30, P) * C).astype(np.int64), C - 1)
        lab = np.full((P, C), 0.1 / max(C - 1, 1), dtype=np.float32)
        lab[np.arange(P), cls] = 0.9
    perm = np.argsort(uniform01(seed, 40, P), kind="stable")
    src_idx = perm[:N]
    tgt_idx = np.concatenate([src_idx[:n_ov], perm[N:N + (M - n_ov)]])
    tgt_idx = tgt_idx[np.argsort(uniform01(seed, 41, M), kind="stable")]

    def cloud(idx, stream):
        n = len(idx)
        xyz = base[idx] + 0.01 * np.stack([normal(seed, stream + k, n) for k in range(3)], axis=1)
        out = {"xyz": xyz.astype(np.float32), "features": None, "labels": None, "geotype": None}
        if feat is not None:
            f = feat[idx] + 0.02 * np.stack([normal(seed, stream + 10 + k, n) for k in range(F)], axis=1)
            out["features"] = f.astype(np.float32)
        if lab is not None:
            out["labels"] = lab[idx].copy()
        if with_geotype:
            g = np.zeros((n, 2), np.float32)
            edge = uniform01(seed, stream + 30, n) < 0.3
            g[edge, 0] = 1.0
            g[~edge, 1] = 1.0
            out["geotype"] = g
        return out

    source = cloud(src_idx, 100)
    target = cloud(tgt_idx, 200)
    Tgt = gt_transform()
    Tinv = np.linalg.inv(Tgt)
    xyz = target["xyz"].astype(np.float64)
    target["xyz"] = (xyz @ Tinv[:3, :3].T + Tinv[:3, 3]).astype(np.float32)
    return {"source": source, "target": target, "T_gt": Tgt}


# ---- the named configurations of SURVEY.md §8(d) / BASELINE.json -----------------
CONFIGS = {
    # name: (P, N, M, seed, F, C)
    "C2": (12_500, 10_000, 10_000, 20_002, 0, 0),
    "C4": (250_000, 200_000, 200_000, 20_004, 5, 0),
    "KITTI05": (20_480, 16_384, 16_384, 20_005, 5, 0),
    "C5": (16_000, 12_800, 12_800, 20_006, 5, 20),
}


def make_config(name: str):
    P, N, M, seed, F, C = CONFIGS[name]
    return make_pair(P, N, M, seed, F=F, C=C)
