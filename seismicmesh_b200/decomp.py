"""`SeismicMesh.decomp` (decomp/blocker.py:4-111): 1-D slab decomposition of a point cloud, the
reference's only parallel strategy.  Host bookkeeping done once per run (not on the hot path)."""
import numpy as np

__all__ = ["blocker"]

# the reference's `axis` convention: 1 cuts along x (dimension 0), 0 along y, 2 along z
_CUT_DIM = {0: 1, 1: 0, 2: 2}


def blocker(points, rank, num_blocks, axis=0):
    """Cut `points` (N, dim) into `num_blocks` equal-width slabs of their bounding box.
    Returns (blocks, block_extents): the points of every non-empty slab and its bounding box
    [min..., max...].  A point exactly on a cut belongs to both neighbours, as in the reference
    (closed intervals, blocker.py:75-100)."""
    points = np.asarray(points)
    num_points, dim = points.shape
    if dim < 2 or dim > 3:
        raise ValueError("Dimensions of points are not supported")
    assert num_points // num_blocks > 1, "too few points for chosen num_blocks"
    if axis not in _CUT_DIM or _CUT_DIM[axis] >= dim:
        raise ValueError("`axis` not supported for points of this dimension")
    d = _CUT_DIM[axis]
    eps = np.finfo(float).eps
    lo, hi = points[:, d].min() - eps, points[:, d].max() + eps
    lows = np.linspace(lo, hi, num_blocks, endpoint=False)
    width = (hi - lo) / num_blocks
    blocks, extents = [], []
    for low in lows:
        block = points[(points[:, d] >= low) & (points[:, d] <= low + width)]
        if block.shape[0]:
            blocks.append(block)
            extents.append([*np.amin(block, axis=0), *np.amax(block, axis=0)])
    return blocks, extents
