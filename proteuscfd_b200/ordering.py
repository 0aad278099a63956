"""Node renumbering helpers used to prepare meshes (host side, numpy).

The reference's SGS (ucs/crs.tcc:90-145) is a sequential Gauss-Seidel in node
order.  If the node numbering is colour-sorted (all nodes of colour 0 first,
then colour 1, ...) the sequential sweep and the parallel multicolour sweep are
the same iteration, so the B200 path can run one colour per kernel launch and
still match the reference bit for bit.
"""
import numpy as np


def node_adjacency(nnode, elems):
    """CSR adjacency (sorted, unique) from an element->node table (any simplex)."""
    elems = np.asarray(elems, dtype=np.int64)
    k = elems.shape[1]
    a, b = [], []
    for i in range(k):
        for j in range(k):
            if i != j:
                a.append(elems[:, i])
                b.append(elems[:, j])
    a = np.concatenate(a)
    b = np.concatenate(b)
    key = np.unique(a * nnode + b)
    a = key // nnode
    b = key % nnode
    ptr = np.zeros(nnode + 1, dtype=np.int64)
    np.add.at(ptr, a + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr, b


def greedy_colors(nnode, ptr, adj):
    """First-fit greedy colouring in natural node order."""
    color = np.full(nnode, -1, dtype=np.int32)
    for i in range(nnode):
        used = color[adj[ptr[i]:ptr[i + 1]]]
        used = used[used >= 0]
        c = 0
        if used.size:
            mask = np.zeros(used.max() + 2, dtype=bool)
            mask[used] = True
            c = int(np.argmin(mask))
        color[i] = c
    return color


def kuhn_box_colors(n):
    """Closed-form 8-colouring of the Kuhn box: colour = (i + 2j + 4k) mod 8.

    Every Kuhn edge joins nodes whose (i,j,k) differ by a non-zero 0/1 vector,
    so the weighted sum differs by 1..7 and never by a multiple of 8.
    """
    np1 = n + 1
    idx = np.arange(np1 ** 3, dtype=np.int64)
    i, j, k = idx % np1, (idx // np1) % np1, idx // (np1 * np1)
    return ((i + 2 * j + 4 * k) % 8).astype(np.int32)


def color_order(color):
    """new id of each old node when nodes are stably sorted by colour."""
    old_of_new = np.argsort(color, kind="stable")
    new_of_old = np.empty_like(old_of_new)
    new_of_old[old_of_new] = np.arange(len(color))
    return new_of_old


def greedy_color_order(nnode, elems):
    ptr, adj = node_adjacency(nnode, elems)
    color = greedy_colors(nnode, ptr, adj)
    return color_order(color), color


def kuhn_box_brick_order(n, brick=(4, 4, 8)):
    """new id of each node of the (n+1)^3 Kuhn-box lattice when the nodes are numbered brick by brick (bricks of
    brick[0] x brick[1] x brick[2] nodes in lexicographic order, nodes inside a brick lexicographic): a block of 128
    consecutive nodes is then a compact brick with ~360 distinct neighbours instead of a 119-node grid line with ~1000,
    which is what the neighbour-row gathers of the gradient / limiter / residual kernels want from L1.  A locality
    ordering in the role the reference gives RCM (ucs/solutionSpace.tcc:61-74); any numbering reproduces the reference's
    arithmetic on that numbering."""
    np1 = n + 1
    idx = np.arange(np1 ** 3, dtype=np.int64)
    i, j, k = idx % np1, (idx // np1) % np1, idx // (np1 * np1)
    bx, by, bz = brick
    nbx, nby = -(-np1 // bx), -(-np1 // by)
    bid = (i // bx) + nbx * ((j // by) + nby * (k // bz))
    loc = (i % bx) + bx * ((j % by) + by * (k % bz))
    old_of_new = np.lexsort((loc, bid))
    new_of_old = np.empty_like(old_of_new)
    new_of_old[old_of_new] = np.arange(len(idx))
    return new_of_old
