"""Deterministic block-world maps shared by make_golden.py and the tests (no reference import)."""
import numpy as np


def build_block_maps(seed, S0, S1, S2, F_sem, F_feat, shift):
    """Deterministic block world shared by the golden generator and the tests
    (tests import this function): a handful of boxes per class written into a
    semantic map and a feature map with plain uniform numbers."""
    rng = np.random.default_rng(seed)
    sem = np.zeros((S0, S1, S2, F_sem), np.float32)
    feat = (rng.random((S0, S1, S2, F_feat)) * 0.01).astype(np.float32) if F_feat else None
    specs = []
    for cls in (3, 7, 7, 7, 12, 12, 45, 45, 50, 20, 20, 20, 20):
        h, w, d = rng.integers(2, 6), rng.integers(2, 6), rng.integers(1, 4)
        y, x, z = rng.integers(0, S0 - 6), rng.integers(0, S1 - 6), rng.integers(0, S2 - 4)
        specs.append((cls, y, x, z, h, w, d))
    for k, (cls, y, x, z, h, w, d) in enumerate(specs):
        if shift and k % 3 == 0:
            y, x = min(y + 5, S0 - h), max(x - 4, 0)
        sem[y:y + h, x:x + w, z:z + d, cls] = (0.2 + 0.8 * rng.random((h, w, d))).astype(np.float32)
        if feat is not None:
            feat[y:y + h, x:x + w, z:z + d] += (0.1 * k + rng.random((h, w, d, F_feat))).astype(np.float32)
    return sem, feat
