"""Flatten a duck-typed (ete3-like) tree + species->count dict into the integer arrays the
device kernels walk.

Node numbering is the reference's default traversal order (ete3 level-order, the row order of
its ancestral-states CSV, AR:1099), so index 0 is the root.  Children keep Newick file order
(AR:221).  Levels are heights above the leaves, so every level only depends on lower ones and
can be one kernel launch.
"""
from __future__ import annotations

import os

import numpy as np

LEAF_MISSING = -1    # species absent from the dict / count None -> uniform vector (AR:209-211)
INTERNAL = -2


class FlatTree:
    """Host-side arrays describing one (tree, data, max_chr) triple."""

    def __init__(self, tree, data, max_chr, schedule=None):
        if schedule is None:
            schedule = os.environ.get("CEVB_SCHED", "child")
        if schedule not in ("child", "parent"):
            raise ValueError("schedule must be 'child' or 'parent'")
        root = tree.get_tree_root()
        nodes = list(root.traverse())            # level-order
        index = {id(nd): k for k, nd in enumerate(nodes)}
        n_nodes = len(nodes)
        self.nodes = nodes
        self.index = index
        self.n_nodes = n_nodes
        self.max_chr = int(max_chr)
        self.n_states = int(max_chr) + 1
        self.root = 0

        parent = np.full(n_nodes, -1, dtype=np.int32)
        child_off = np.zeros(n_nodes + 1, dtype=np.int32)
        child_idx = []
        for k, nd in enumerate(nodes):
            for ch in nd.children:
                child_idx.append(index[id(ch)])
                parent[index[id(ch)]] = k
            child_off[k + 1] = len(child_idx)
        self.parent = parent
        self.child_off = child_off
        self.child_idx = np.asarray(child_idx, dtype=np.int32)

        # raw branch lengths (for the simulator, AR:562-566) and the likelihood's effective
        # lengths (None / negative -> 1e-6, AR:224-227)
        raw = np.zeros(n_nodes, dtype=np.float64)
        eff = np.zeros(n_nodes, dtype=np.float64)
        for k, nd in enumerate(nodes):
            d = getattr(nd, "dist", None)
            raw[k] = np.nan if d is None else float(d)
            eff[k] = 1e-6 if (d is None or d < 0) else float(d)
        self.dist_raw = raw
        self.dist_eff = eff

        # one transition matrix per distinct effective length among the non-root nodes
        non_root = np.arange(1, n_nodes)
        uniq, inverse = np.unique(eff[non_root], return_inverse=True)
        if uniq.size and np.isnan(uniq).any():     # np.unique keeps NaNs; map each to itself
            uniq, inverse = eff[non_root].copy(), np.arange(non_root.size)
        self.t_unique = np.ascontiguousarray(uniq, dtype=np.float64)
        branch_of = np.zeros(n_nodes, dtype=np.int32)
        branch_of[non_root] = inverse.astype(np.int32)
        self.branch_of = branch_of
        self.n_branches = n_nodes - 1
        self.n_t = int(self.t_unique.size)

        # leaf observation codes (AR:201-214)
        leaf_code = np.full(n_nodes, INTERNAL, dtype=np.int32)
        n_data = 0
        for k, nd in enumerate(nodes):
            if nd.children:
                continue
            obs = data.get(nd.name) if data is not None else None
            if obs is None:
                leaf_code[k] = LEAF_MISSING
            else:
                n_data += 1
                inside = False
                try:
                    inside = 0 <= obs <= max_chr
                except TypeError:
                    inside = False
                leaf_code[k] = int(obs) if inside else int(max_chr)
        self.leaf_code = leaf_code
        self.leaf_names = [nd.name for nd in nodes if not nd.children]
        self.n_leaves = int(np.sum(leaf_code != INTERNAL))
        self.n_leaves_with_data = n_data

        # levels by height; internal nodes only
        height = np.zeros(n_nodes, dtype=np.int32)
        for k in range(n_nodes - 1, -1, -1):       # reverse level-order: children before parents
            lo, hi = child_off[k], child_off[k + 1]
            if hi > lo:
                height[k] = 1 + max(height[c] for c in self.child_idx[lo:hi])
        self.height = height
        n_levels = int(height.max()) if n_nodes else 0
        order = [k for k in np.argsort(height, kind="stable") if height[k] > 0]
        self.level_nodes = np.asarray(order, dtype=np.int32)
        counts = np.bincount(height[height > 0], minlength=n_levels + 1)[1:]
        level_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.level_off = level_off
        self.n_levels = n_levels

        # Work lists of the fused kernels.  Branches above observed leaves depend on nothing
        # below them, so all of them (every level) form one list; the remaining branches
        # (internal or missing-data children) are grouped by the HEIGHT OF THE CHILD: list lv
        # is launched right after the nodes of height lv are combined, i.e. as early as the
        # branch's input exists (a parent at level lv only has children of height <= lv).  That
        # moves work out of the narrow launches near the root into the wide early ones
        # (1,000-taxon bench tree: 343/196/124/... branches per launch instead of 237/176/139/...).
        # schedule="parent" keeps the round-1 grouping by the level of the PARENT (A/B runs).
        # Longest first inside each list, so that the 8-branch groups of the tensor-core kernel
        # need similar series lengths.
        tkey = lambda c: -(eff[c] if eff[c] == eff[c] else 0.0)
        self.leaf_child = np.asarray(sorted((c for c in range(1, n_nodes) if leaf_code[c] >= 0),
                                            key=tkey), dtype=np.int32)
        lvl_child = []
        lvl_child_off = [0]
        self.schedule = schedule
        for lv in range(n_levels):
            if schedule == "parent":
                kids = []
                for p in self.level_nodes[level_off[lv]:level_off[lv + 1]]:
                    kids.extend(self.child_idx[child_off[p]:child_off[p + 1]].tolist())
            else:
                kids = [c for c in range(1, n_nodes) if height[c] == lv]
            lvl_child.extend(sorted((c for c in kids if leaf_code[c] < 0), key=tkey))
            lvl_child_off.append(len(lvl_child))
        self.level_child = np.asarray(lvl_child, dtype=np.int32)
        self.level_child_off = np.asarray(lvl_child_off, dtype=np.int32)

        # depth below the root: node ids are level-order, so every depth is a contiguous id range
        # (the marginal down-pass and the joint backtrack walk the tree one depth per step)
        depth = np.zeros(n_nodes, dtype=np.int32)
        for k in range(1, n_nodes):
            depth[k] = depth[parent[k]] + 1
        assert np.all(np.diff(depth) >= 0), "level-order numbering expected"
        self.depth = depth
        self.depth_off = np.concatenate([[0], np.cumsum(np.bincount(depth))]).astype(np.int32)
        self.n_depths = int(self.depth_off.size - 1)

        # pre-order (parents before children, children in file order) for the simulator (AR:551)
        pre = []
        stack = [0]
        while stack:
            k = stack.pop()
            pre.append(k)
            lo, hi = child_off[k], child_off[k + 1]
            stack.extend(reversed(self.child_idx[lo:hi].tolist()))
        self.preorder = np.asarray(pre, dtype=np.int32)

    def signature(self):
        return (self.n_nodes, self.n_states, self.n_t)

    def algorithmic_prune_bytes(self):
        """Bytes one pruning pass must read/write per parameter vector:
        a full P per branch whose child is internal or has missing data, one column of P per
        observed-leaf branch, plus the node vectors."""
        n = self.n_states
        full = 0
        col = 0
        for k in range(1, self.n_nodes):
            if self.leaf_code[k] >= 0:
                col += 1
            else:
                full += 1
        return 8 * (full * (n * n + n) + col * n + self.n_nodes * n)

