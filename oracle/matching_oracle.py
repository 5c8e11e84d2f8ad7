"""CPU ORACLE (test infrastructure, not product code) -- descriptor matching.

Restates util/uio.py:245-258 (`knn_search`: exact 1-NN under L2, one KD-tree query per row) and the mutual check of
scripts/evaluation_3dmatch.py:207-217 with a float64 brute force (an exact KD-tree query and a brute-force argmin return the
same index unless two candidates are exactly equidistant).  Parity: not pinned by a reference-held vector (the reference has
none for this step); validated against scipy.spatial.cKDTree, which lib/eval.py:10-16 (`find_nn_cpu`) uses.
Only tests/ may import this."""
import numpy as np


def knn_search(points_src, points_dst):
    src, dst = np.asarray(points_src, np.float64), np.asarray(points_dst, np.float64)
    out = np.empty(len(src), np.int32)
    for s in range(0, len(src), 512):
        d = ((src[s:s + 512, None, :] - dst[None, :, :]) ** 2).sum(-1)
        out[s:s + 512] = d.argmin(1)
    return out


def mutual(frag1_descs, frag2_descs):
    nn21 = knn_search(frag2_descs, frag1_descs)
    nn12 = knn_search(frag1_descs, frag2_descs)
    return nn21, np.flatnonzero(np.equal(np.arange(len(nn21)), nn12[nn21]))
