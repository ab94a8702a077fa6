"""ete3-free Newick tree with the slice of the ete3 node API the hot path touches.

The reference walks an ``ete3.Tree`` (reference ``src/ancestral_reconstruction.py:2``,
``:990`` ``Tree(tree_file, format=1)``).  ete3 is not installable here, and the likelihood
path only needs a small duck-typed surface: ``children`` in file
order, ``up``, ``dist``, ``name``, ``is_leaf``/``is_root``/``get_tree_root``,
``traverse`` (default level-order, plus pre/post-order), ``add_features``, ``get_leaves``,
iteration over leaves and ``get_ascii``.  Format-1 semantics are kept: internal labels are
node names, a missing ``:dist`` is 1.0 and a missing name is ``""``.

Any object with the same attributes (a real ``ete3.Tree`` included) is accepted by the
rest of the package; nothing below is required to be *this* class.
"""
from __future__ import annotations

import os
from collections import deque

DEFAULT_DIST = 1.0
DEFAULT_NAME = ""


class NewickError(ValueError):
    pass


class TreeNode:
    """One node of a rooted tree; the root object doubles as "the tree"."""

    def __init__(self, newick=None, format=1, name=None, dist=None):
        self.children = []
        self.up = None
        self.dist = DEFAULT_DIST if dist is None else dist
        self.name = DEFAULT_NAME if name is None else name
        self.support = 1.0
        self.features = {"dist", "name", "support"}
        if newick is not None:
            _read_newick(newick, self, format)

    # -- structure ---------------------------------------------------------
    def add_child(self, child=None, name=None, dist=None):
        if child is None:
            child = self.__class__(name=name, dist=dist)
        else:
            if name is not None:
                child.name = name
            if dist is not None:
                child.dist = dist
        child.up = self
        self.children.append(child)
        return child

    def is_leaf(self):
        return len(self.children) == 0

    def is_root(self):
        return self.up is None

    def get_tree_root(self):
        node = self
        while node.up is not None:
            node = node.up
        return node

    # -- features ----------------------------------------------------------
    def add_feature(self, key, value):
        setattr(self, key, value)
        self.features.add(key)

    def add_features(self, **features):
        for key, value in features.items():
            setattr(self, key, value)
            self.features.add(key)

    # -- traversal ---------------------------------------------------------
    def traverse(self, strategy="levelorder", is_leaf_fn=None):
        if strategy == "levelorder":
            return self._iter_levelorder(is_leaf_fn)
        if strategy == "preorder":
            return self._iter_preorder(is_leaf_fn)
        if strategy == "postorder":
            return self._iter_postorder(is_leaf_fn)
        raise ValueError(f"unknown traversal strategy {strategy!r}")

    def _iter_levelorder(self, is_leaf_fn=None):
        queue = deque([self])
        while queue:
            node = queue.popleft()
            yield node
            if is_leaf_fn is None or not is_leaf_fn(node):
                queue.extend(node.children)

    def _iter_preorder(self, is_leaf_fn=None):
        stack = [self]
        while stack:
            node = stack.pop()
            yield node
            if is_leaf_fn is None or not is_leaf_fn(node):
                stack.extend(reversed(node.children))

    def _iter_postorder(self, is_leaf_fn=None):
        # iterative: (node, visited-children?) pairs, so deep trees do not recurse
        stack = [(self, False)]
        while stack:
            node, expanded = stack.pop()
            if expanded or not node.children or (is_leaf_fn is not None and is_leaf_fn(node)):
                yield node
            else:
                stack.append((node, True))
                for child in reversed(node.children):
                    stack.append((child, False))

    def iter_leaves(self, is_leaf_fn=None):
        for node in self._iter_preorder(is_leaf_fn):
            if (is_leaf_fn(node) if is_leaf_fn is not None else node.is_leaf()):
                yield node

    def get_leaves(self, is_leaf_fn=None):
        return list(self.iter_leaves(is_leaf_fn))

    def get_leaf_names(self):
        return [leaf.name for leaf in self.iter_leaves()]

    def get_descendants(self, strategy="levelorder"):
        return [n for n in self.traverse(strategy) if n is not self]

    def __iter__(self):
        return self.iter_leaves()

    def __len__(self):
        return sum(1 for _ in self.iter_leaves())

    def __bool__(self):
        return True

    def __repr__(self):
        return f"Tree node '{self.name}' ({hex(id(self))})"

    # -- text output -------------------------------------------------------
    def get_ascii(self, show_internal=True, compact=False, attributes=None):
        """Indented text rendering (reference logs this at AR:1093; layout is free-form)."""
        attributes = attributes or ["name"]
        lines = []
        stack = [(self, 0)]
        while stack:
            node, depth = stack.pop()
            if node.is_leaf() or show_internal:
                label = ", ".join(str(getattr(node, a)) for a in attributes if hasattr(node, a))
            else:
                label = ""
            lines.append("   " * depth + ("\\-" if depth else "") + label)
            for child in reversed(node.children):
                stack.append((child, depth + 1))
        return "\n" + "\n".join(lines)

    def write(self, format=1, dist_formatter="%0.6g"):
        """Serialise back to Newick (names on all nodes, lengths on all but the root)."""
        out = []
        # iterative post-order assembly
        stack = [(self, 0)]
        pieces = {}
        for node in self._iter_postorder():
            label = _quote_if_needed(node.name)
            if node.children:
                text = "(" + ",".join(pieces.pop(id(c)) for c in node.children) + ")" + label
            else:
                text = label
            if node.up is not None:
                text += ":" + (dist_formatter % node.dist)
            pieces[id(node)] = text
        del stack, out
        return pieces[id(self)] + ";"


Tree = TreeNode


def _quote_if_needed(name):
    if any(ch in name for ch in "(),:;[] \t'"):
        return "'" + name.replace("'", "''") + "'"
    return name


def _read_newick(source, root, format=1):
    if format not in (0, 1, 2, 3, 5, 100):
        raise NewickError(f"unsupported newick format {format}")
    if isinstance(source, (str, os.PathLike)) and os.path.exists(os.fspath(source)):
        with open(source, "r") as handle:
            text = handle.read()
    else:
        text = os.fspath(source) if isinstance(source, os.PathLike) else str(source)
    text = text.strip()
    if not text:
        raise NewickError("empty newick string")
    end = _find_terminator(text)
    if end < 0:
        raise NewickError("newick string does not end with ';'")
    text = text[:end]
    _parse(text, root, format)


def _find_terminator(text):
    in_quote = False
    depth_comment = 0
    last = -1
    for i, ch in enumerate(text):
        if in_quote:
            if ch == "'":
                in_quote = False
            continue
        if depth_comment:
            if ch == "]":
                depth_comment -= 1
            elif ch == "[":
                depth_comment += 1
            continue
        if ch == "'":
            in_quote = True
        elif ch == "[":
            depth_comment += 1
        elif ch == ";":
            last = i
            break
    return last


def _parse(text, root, fmt):
    """Single left-to-right pass; ``current`` is the node whose label is being read."""
    current = root
    i = 0
    n = len(text)
    expecting_label = True  # a fresh node may receive name / :dist next
    token_name = []
    token_dist = []
    reading_dist = False

    def flush(node):
        nonlocal token_name, token_dist, reading_dist
        name = "".join(token_name).strip()
        dist = "".join(token_dist).strip()
        if name:
            if fmt in (0, 2) and node.children:
                try:
                    node.support = float(name)
                except ValueError:
                    node.name = name
            else:
                node.name = name
        if dist:
            try:
                node.dist = float(dist)
            except ValueError as exc:
                raise NewickError(f"bad branch length {dist!r}") from exc
        token_name, token_dist, reading_dist = [], [], False

    while i < n:
        ch = text[i]
        if ch == "'":
            j = i + 1
            buf = []
            while j < n:
                if text[j] == "'":
                    if j + 1 < n and text[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(text[j])
                j += 1
            if j >= n:
                raise NewickError("unterminated quoted label")
            (token_dist if reading_dist else token_name).append("".join(buf))
            i = j + 1
            continue
        if ch == "[":
            depth = 1
            j = i + 1
            while j < n and depth:
                if text[j] == "[":
                    depth += 1
                elif text[j] == "]":
                    depth -= 1
                j += 1
            if depth:
                raise NewickError("unterminated comment")
            i = j
            continue
        if ch == "(":
            if token_name or token_dist:
                raise NewickError("unexpected '(' after a label")
            current = current.add_child()
            expecting_label = True
        elif ch == ",":
            flush(current)
            if current.up is None:
                raise NewickError("',' outside of parentheses")
            current = current.up.add_child()
            expecting_label = True
        elif ch == ")":
            flush(current)
            if current.up is None:
                raise NewickError("unbalanced ')'")
            current = current.up
            expecting_label = True
        elif ch == ":":
            reading_dist = True
        elif ch in " \t\r\n" and not reading_dist and not token_name:
            pass
        else:
            (token_dist if reading_dist else token_name).append(ch)
        i += 1
    flush(current)
    if current is not root:
        raise NewickError("unbalanced '(' in newick string")
    del expecting_label
