"""CPU: property tests (hypothesis) of the host-side flattening that every device kernel walks:
random topologies with polytomies, missing / out-of-range counts and odd branch lengths."""
import numpy as np
from hypothesis import given, settings, strategies as st

from conftest import make_tree


def _random_newick(rng, n_leaves):
    """Random rooted tree with polytomies; names L0.., some lengths missing / zero / negative."""
    nodes = [f"L{k}" for k in range(n_leaves)]

    def length():
        r = rng.random()
        if r < 0.08:
            return ""                                   # no length -> ete3 default 1.0
        if r < 0.14:
            return ":0"
        if r < 0.18:
            return ":-0.5"                              # negative -> 1e-6 in the likelihood (AR:224-227)
        return f":{rng.choice([0.01, 0.05, 0.05, 0.2, 1.5]) * (1 + rng.integers(0, 3)):.6g}"

    items = [nm + length() for nm in nodes]
    k = 0
    while len(items) > 1:
        take = int(min(len(items), rng.choice([2, 2, 2, 3, 4])))
        idx = sorted(rng.choice(len(items), size=take, replace=False).tolist(), reverse=True)
        kids = [items.pop(i) for i in idx]
        name = f"I{k}" if rng.random() < 0.5 else ""
        k += 1
        sub = "(" + ",".join(kids) + ")" + name
        items.append(sub + (length() if len(items) > 0 else ""))
    return items[0] + ";"


@settings(max_examples=60, deadline=None)
@given(st.integers(min_value=0, max_value=2 ** 31 - 1), st.integers(min_value=2, max_value=40))
def test_flat_tree_invariants(seed, n_leaves):
    from chromevol_enhanced_b200.flatten import FlatTree, INTERNAL, LEAF_MISSING
    rng = np.random.default_rng(seed)
    tree = make_tree(_random_newick(rng, n_leaves))
    max_chr = 20
    data = {}
    for k in range(n_leaves):
        r = rng.random()
        if r < 0.15:
            continue                                    # species absent -> uniform vector
        if r < 0.2:
            data[f"L{k}"] = None
        elif r < 0.3:
            data[f"L{k}"] = int(rng.choice([-3, 21, 400]))   # out of range -> clamped to max (AR:204-208)
        else:
            data[f"L{k}"] = int(rng.integers(0, max_chr + 1))
    f = FlatTree(tree, data, max_chr)
    nodes = list(tree.traverse())
    assert f.n_nodes == len(nodes) and f.root == 0 and f.nodes[0] is tree.get_tree_root()
    # children in file order, parents consistent
    for k, nd in enumerate(f.nodes):
        kids = f.child_idx[f.child_off[k]:f.child_off[k + 1]]
        assert [f.nodes[c] for c in kids] == list(nd.children)
        assert all(f.parent[c] == k for c in kids)
    assert f.parent[0] == -1 and (f.parent[1:] >= 0).all()
    # leaf codes
    for k, nd in enumerate(f.nodes):
        if nd.children:
            assert f.leaf_code[k] == INTERNAL
        else:
            obs = data.get(nd.name)
            if obs is None:
                assert f.leaf_code[k] == LEAF_MISSING
            else:
                assert f.leaf_code[k] == (obs if 0 <= obs <= max_chr else max_chr)
    # levels: heights, every internal node exactly once, children strictly below their parent
    internal = [k for k, nd in enumerate(f.nodes) if nd.children]
    assert sorted(f.level_nodes.tolist()) == internal
    assert f.level_off[0] == 0 and f.level_off[-1] == len(internal) and len(f.level_off) == f.n_levels + 1
    level_of = {}
    for lv in range(f.n_levels):
        for p in f.level_nodes[f.level_off[lv]:f.level_off[lv + 1]]:
            level_of[int(p)] = lv
    for p in internal:
        kids = f.child_idx[f.child_off[p]:f.child_off[p + 1]]
        below = [level_of[int(c)] for c in kids if int(c) in level_of]
        assert level_of[p] == (max(below) + 1 if below else 0)
    # effective lengths and the unique-length table
    for k, nd in enumerate(f.nodes):
        d = nd.dist
        assert f.dist_eff[k] == (1e-6 if (d is None or d < 0) else float(d))
    for k in range(1, f.n_nodes):
        assert f.t_unique[f.branch_of[k]] == f.dist_eff[k]
    assert len(np.unique(f.t_unique)) == f.n_t
    # work lists of the fused kernels partition the branches
    leaf_list = f.leaf_child.tolist()
    level_list = f.level_child.tolist()
    assert sorted(leaf_list + level_list) == list(range(1, f.n_nodes))
    assert all(f.leaf_code[c] >= 0 for c in leaf_list) and all(f.leaf_code[c] < 0 for c in level_list)
    assert all(f.dist_eff[a] >= f.dist_eff[b] for a, b in zip(leaf_list, leaf_list[1:]))
    for lv in range(f.n_levels):
        part = f.level_child[f.level_child_off[lv]:f.level_child_off[lv + 1]].tolist()
        # a branch is launched as soon as its child's vector exists (child height == list index),
        # never after its parent is combined
        assert all(f.height[c] == lv and level_of[int(f.parent[c])] >= lv for c in part)
        assert all(f.dist_eff[a] >= f.dist_eff[b] for a, b in zip(part, part[1:]))
    # depths below the root: level-order ids make every depth a contiguous id range (the marginal
    # down-pass and the joint backtrack launch one step per depth)
    assert f.depth[0] == 0 and f.depth_off[0] == 0 and f.depth_off[-1] == f.n_nodes
    assert f.n_depths == int(f.depth.max()) + 1 == len(f.depth_off) - 1
    for k in range(1, f.n_nodes):
        assert f.depth[k] == f.depth[f.parent[k]] + 1
    for d in range(f.n_depths):
        assert (f.depth[f.depth_off[d]:f.depth_off[d + 1]] == d).all()
    # pre-order for the simulator: parents before children, root first
    pos = {int(v): i for i, v in enumerate(f.preorder.tolist())}
    assert f.preorder[0] == 0 and sorted(pos) == list(range(f.n_nodes))
    assert all(pos[int(f.parent[k])] < pos[k] for k in range(1, f.n_nodes))
