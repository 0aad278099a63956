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


# ---- the reference's node reordering (Mesh::ReorderMeshCuthillMcKee, ucs/mesh.tcc:2412-2494; the default of ucs.x:
# reorderMesh = 1, solutionSpace.tcc:61-74).  Set-up, on the host.
def _std_sort(keys, items):
    """libstdc++'s std::sort (introsort + final insertion sort, threshold 16) on `items` ordered by `keys` with a
    strict-less comparison, reproduced step by step: the reference sorts the candidates of a front by degree ONLY with
    the unstable std::sort (mesh.tcc:2459, oddsNends.cpp:4-7), so the order of equal degrees -- and with it the whole
    permutation -- is whatever this algorithm leaves.  Returns the sorted items."""
    k, v = list(keys), list(items)
    n = len(k)

    def swap(i, j):
        k[i], k[j] = k[j], k[i]
        v[i], v[j] = v[j], v[i]

    def unguarded_linear_insert(last):
        kv, vv = k[last], v[last]
        nxt = last - 1
        while kv < k[nxt]:
            k[last], v[last] = k[nxt], v[nxt]
            last = nxt
            nxt -= 1
        k[last], v[last] = kv, vv

    def insertion_sort(first, last):
        for i in range(first + 1, last):
            if k[i] < k[first]:
                kv, vv = k[i], v[i]
                k[first + 1:i + 1] = k[first:i]
                v[first + 1:i + 1] = v[first:i]
                k[first], v[first] = kv, vv
            else:
                unguarded_linear_insert(i)

    def introsort_loop(first, last, depth):
        while last - first > 16:
            if depth == 0:
                raise NotImplementedError("std::sort fell back to heapsort: not reproduced (never for valence-sized fronts)")
            depth -= 1
            mid = first + (last - first) // 2
            a, b, c = first + 1, mid, last - 1
            if k[a] < k[b]:
                if k[b] < k[c]:
                    swap(first, b)
                elif k[a] < k[c]:
                    swap(first, c)
                else:
                    swap(first, a)
            elif k[a] < k[c]:
                swap(first, a)
            elif k[b] < k[c]:
                swap(first, c)
            else:
                swap(first, b)
            lo, hi = first + 1, last
            while True:
                while k[lo] < k[first]:
                    lo += 1
                hi -= 1
                while k[first] < k[hi]:
                    hi -= 1
                if not lo < hi:
                    break
                swap(lo, hi)
                lo += 1
            introsort_loop(lo, last, depth)
            last = lo

    if n > 1:
        introsort_loop(0, n, 2 * (n.bit_length() - 1))
        if n > 16:
            insertion_sort(0, 16)
            for i in range(16, n):
                unguarded_linear_insert(i)
        else:
            insertion_sort(0, n)
    return v


def cuthill_mckee(nnode, ipsp, psp, reverse=True, startnode=0):
    """ordering[new] = old as the reference computes it: breadth-first from node 0, the unvisited local neighbours of
    the current node appended in psp order and sorted by degree (std::sort, see _std_sort), ghosts (>= nnode) ignored,
    the whole list reversed for `reverse` (the solver's choice, solutionSpace.tcc:65).  A graph that is not connected
    makes the reference spin forever (its re-seeding never marks the new seed, mesh.tcc:2470-2479); here it raises."""
    from collections import deque
    ipsp = np.asarray(ipsp, dtype=np.int64)
    psp = np.asarray(psp, dtype=np.int64)
    degree = ipsp[1:nnode + 1] - ipsp[:nnode]
    used = np.zeros(nnode, dtype=bool)
    ordering = [int(startnode)]
    used[startnode] = True
    queue = deque()
    z = int(startnode)
    while len(ordering) < nnode:
        front = []
        for p in psp[ipsp[z]:ipsp[z + 1]]:
            if p < nnode and not used[p]:
                front.append(int(p))
                used[p] = True
        for p in _std_sort([int(degree[p]) for p in front], front):
            ordering.append(p)
            queue.append(p)
        if not queue:
            if len(ordering) < nnode:
                raise ValueError("cuthill_mckee: the node graph is not connected")
            break
        z = queue.popleft()
    out = np.array(ordering, dtype=np.int32)
    return out[::-1].copy() if reverse else out
