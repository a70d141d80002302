"""Drop-in for `split_into_instances` of `Fusion3DSeg/segUtils/cv.py` (reference `cv.py:402-500`), SURVEY 8(f) rank 1.

The reference flood-fills equal-class neighbours with a Python BFS (`list.pop(0)`, `cv.py:425-440`).  Here the
connected components come from the GPU union-find of `csrc/box_merge.cu` (`f3d_union_find`, root = smallest point
index of the component) over the equal-class edges of the adjacency list; the reference's numbering is then
reproduced from the component table: classes in `instance_classes` order, inside a class by ascending smallest
point index (the BFS seeds are `remaining_points[0]`, `cv.py:473-475`), small components folded into one
"small disjoint" instance created at the position of the first small component (`cv.py:478-486`).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import engine
from ..._lib import require_cuda


def adjacency_to_csr(adj):
    """list / object array of neighbour-index arrays (`fusion.py:374-377`) or an (indptr, indices) pair (numpy arrays or
    device tensors, e.g. straight from `engine.radius_adjacency`) -> CSR int64 device tensors."""
    dev = require_cuda()
    if isinstance(adj, tuple) and len(adj) == 2:
        ip, ix = adj
        if isinstance(ip, torch.Tensor):
            return ip.to(dev, torch.int64), ix.to(dev, torch.int64)
        return torch.as_tensor(np.asarray(ip, dtype=np.int64)).to(dev), torch.as_tensor(np.asarray(ix, dtype=np.int64)).to(dev)
    counts = np.fromiter((len(a) for a in adj), dtype=np.int64, count=len(adj))
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    indices = np.concatenate([np.asarray(a, dtype=np.int64) for a in adj]) if len(adj) else np.zeros(0, np.int64)
    return torch.as_tensor(indptr).to(dev), torch.as_tensor(indices).to(dev)


def connected_components(classes, adj):
    """labels int64 [N] (device): smallest point index of the point's equal-class connected component.

    The adjacency is treated as an UNDIRECTED graph: every listed pair (i, j), i != j, links i and j whichever row lists
    it.  For the symmetric lists the reference produces (`KDTree.query_radius`, `fusion.py:374-375`) this is exactly what
    its BFS (`cv.py:425-440`) computes; a hand-made one-directional list is closed symmetrically here, whereas the
    reference BFS would follow it as directed."""
    dev = require_cuda()
    indptr, dst = adjacency_to_csr(adj)
    n = int(indptr.shape[0]) - 1
    cls = torch.as_tensor(np.ascontiguousarray(classes)).to(dev)
    src = torch.repeat_interleave(torch.arange(n, device=dev), indptr[1:] - indptr[:-1])
    keep = (cls[src] == cls[dst]) & (src != dst)
    a, b = src[keep], dst[keep]
    edges = torch.stack([torch.minimum(a, b), torch.maximum(a, b)], dim=1).to(torch.int32).contiguous()
    return engine.union_find(n, edges).to(torch.int64)


def split_into_instances(classes, adj, nclasses=133, instance_classes=None, minimum_points=1, verbose=False):
    """Same contract as the reference (`cv.py:402-423`): returns (instance ids [M], point instance ids [N],
    info list of {'id','isthing','category_id','area'}, updated point classes [N])."""
    classes = np.asarray(classes).copy()
    roots_of = connected_components(classes, adj).cpu().numpy()
    allclasses = np.unique(classes)
    ids = np.zeros_like(classes)
    info = []
    small_id = None
    if instance_classes is None:                                   # cv.py:448-456
        inst = allclasses
        ninst, sem = 0, np.zeros(0, dtype=classes.dtype)
        if (inst == nclasses).any():
            inst = inst[inst != nclasses]
            sem, ninst = np.array([nclasses]), 1
    else:                                                          # cv.py:457-460
        inst = np.array(instance_classes)
        sem = np.setdiff1d(allclasses, inst)
        ninst = len(sem)
    for k in range(ninst if len(sem) else 0):                      # cv.py:462-470
        m = classes == sem[k]
        ids[m] = k
        info.append({'id': k, 'isthing': False, 'category_id': int(sem[k]), 'area': int(m.sum())})
        if sem[k] == nclasses:
            small_id = k

    # component table restricted to the instance classes, in the reference's encounter order
    roots, inverse, sizes = np.unique(roots_of, return_inverse=True, return_counts=True)
    rcls = classes[roots]
    order_of_class = {int(c): i for i, c in reversed(list(enumerate(inst)))}   # first occurrence wins
    rank = np.array([order_of_class.get(int(c), -1) for c in rcls])
    sel = np.nonzero(rank >= 0)[0]
    sel = sel[np.lexsort((roots[sel], rank[sel]))]
    small = sizes[sel] < minimum_points
    takes_id = ~small
    if small.any() and small_id is None:
        first = int(np.argmax(small))
        takes_id = takes_id.copy()
        takes_id[first] = True                                      # the "small disjoint" instance is created here
    run = ninst + np.cumsum(takes_id) - 1
    if small.any() and small_id is None:
        small_id = int(run[int(np.argmax(small))])
    comp_id = np.where(small, -1 if small_id is None else small_id, run)
    created = np.nonzero(takes_id)[0]
    for j in created:                                              # info entries in creation order == id order
        if small[j]:
            info.append({'id': int(run[j]), 'isthing': True, 'category_id': int(nclasses), 'area': 0})
        else:
            info.append({'id': int(run[j]), 'isthing': True, 'category_id': int(rcls[sel[j]]), 'area': int(sizes[sel[j]])})
    if small.any():
        info[small_id]['area'] += int(sizes[sel][small].sum())
    ninst += int(takes_id.sum())

    comp_to_id = np.full(len(roots), -1, dtype=np.int64)
    comp_to_id[sel] = comp_id
    comp_small = np.zeros(len(roots), dtype=bool)
    comp_small[sel] = small
    pid = comp_to_id[inverse]
    touched = pid >= 0
    ids[touched] = pid[touched]
    classes[comp_small[inverse]] = nclasses                         # cv.py:478,499
    if verbose:
        print(f'split {len(sel)} components into {ninst} instances')
    return np.arange(ninst), ids, info, classes
