"""Newick -> flat tree in the node order PhyFoo's parser produces.

Host-side mirror of io/NewickToBayesNet.java:38-208 (reference): the root exists before the first
token; "(" opens a child of the current node, "," closes the current node and opens a sibling, ")"
closes it and returns to the parent, ":" announces the branch length of the node being closed.
Nodes therefore come out in DFS pre-order with parents before children, and the leaves in
left-to-right order are the species order (util/Util.java:39-53).  NHX comments, quoted labels and
escapes (:152-193) are not supported: PhyFoo's own inputs never use them.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

_PUNCT = "(),:;"


@dataclass
class Tree:
    parent: np.ndarray          # int32 [N], -1 at the root
    branch: np.ndarray          # float64 [N], distance to parent (0 at the root / when absent)
    leaf_species: np.ndarray    # int32 [N], species index at leaves, -1 at internal nodes
    labels: List[str] = field(default_factory=list)

    @property
    def n_nodes(self):
        return int(self.parent.size)

    @property
    def n_species(self):
        return int((self.leaf_species >= 0).sum())

    @property
    def species(self):
        """Leaf labels in species order."""
        order = np.argsort(np.where(self.leaf_species >= 0, self.leaf_species, 1 << 30))[: self.n_species]
        return [self.labels[k] for k in order]


def parse_newick(text: str) -> Tree:
    parent, branch, labels = [-1], [0.0], [""]
    cur = 0
    label, dist, expect_dist = "", None, False

    def finish(node):
        nonlocal label, dist
        if label:
            labels[node] = label
        if dist is not None:
            branch[node] = dist
        label, dist = "", None

    def new_node(par):
        parent.append(par)
        branch.append(0.0)
        labels.append("")
        return len(parent) - 1

    pos, n = 0, len(text)
    while pos < n:
        ch = text[pos]
        if ord(ch) <= 10:
            pos += 1
        elif ch == "(":
            cur = new_node(cur)
            pos += 1
        elif ch == ",":
            finish(cur)
            if parent[cur] < 0:
                raise ValueError("newick: ',' outside any parenthesis")
            cur = new_node(parent[cur])
            pos += 1
        elif ch == ")":
            finish(cur)
            if parent[cur] < 0:
                raise ValueError("newick: unbalanced ')'")
            cur = parent[cur]
            pos += 1
        elif ch == ":":
            expect_dist = True
            pos += 1
        elif ch == ";":
            break
        else:
            end = pos
            while end < n and text[end] not in _PUNCT and ord(text[end]) > 10:
                end += 1
            word = text[pos:end]
            if expect_dist:
                dist, expect_dist = float(word), False
            else:
                label += word
            pos = end
    N = len(parent)
    has_child = np.zeros(N, bool)
    for k in range(1, N):
        has_child[parent[k]] = True
    leaf_species = np.full(N, -1, np.int32)
    leaf_species[~has_child] = np.arange(int((~has_child).sum()), dtype=np.int32)
    return Tree(np.asarray(parent, np.int32), np.asarray(branch, np.float64), leaf_species, labels)


def species_order(newick: str) -> List[str]:
    """util/Util.java:39-53 getOrderedArrayListFromTree: strip whitespace, distances and brackets, split on ','."""
    import re
    s = re.sub(r"\s", "", newick)
    s = re.sub(r":[\d]*[\.][\d]+", "", s)
    s = re.sub(r":[\d]+", ":", s)
    s = re.sub(r"[\[\]()\;]", "", s)
    return s.split(",")
