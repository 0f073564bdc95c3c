"""Parser for mope's Newick dialect -> flat postorder branch tables.

Semantics follow the reference's modified python-newick (mope/newick.py:42-102,164-214):
`name:var` is a fixed-length branch whose length is the parameter `var`; `name:var*mult` scales the
parameter by the per-family multiplier column `mult` (an age); `name:var^` marks a bottleneck branch
(the parameter is a bottleneck size).  Postorder is children left to right, then the node.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

RESERVED = ":;,()"


@dataclass
class Node:
    name: Optional[str]
    varname: Optional[str] = None
    multipliername: Optional[str] = None
    is_bottleneck: bool = False
    children: List["Node"] = field(default_factory=list)
    parent: Optional["Node"] = None

    @property
    def is_leaf(self) -> bool:
        return not self.children

    def postorder(self):
        # iterative: deep (unary-chain) trees must not hit the recursion limit
        stack = [(self, 0)]
        while stack:
            node, i = stack.pop()
            if i < len(node.children):
                stack.append((node, i + 1))
                stack.append((node.children[i], 0))
            else:
                yield node


def _parse_label(label: str) -> Node:
    name, length = label, None
    if ":" in label:
        name, length = label.split(":", 1)
    name = name.strip() or None
    node = Node(name=name)
    if length is not None and length.strip() != "":
        length = length.strip()
        for ch in RESERVED:
            if ch in length:
                raise ValueError('Node name ({}) or branch length ({}) must not contain "{}"'.format(name, length, ch))
        if "*" in length:
            parts = length.split("*")
            if len(parts) != 2:
                raise ValueError("invalid length in tree file")
            node.varname, node.multipliername = parts[0], parts[1]
        else:
            node.varname = length
        if node.varname.endswith("^"):
            node.is_bottleneck = True
            node.varname = node.varname[:-1]
    return node


def loads(text: str) -> List[Node]:
    """All trees of a Newick string (separated by ';')."""
    return [_parse_tree(part.strip()) for part in text.split(";") if part.strip()]


def _parse_tree(s: str) -> Node:
    pos = 0
    n = len(s)

    def parse() -> Node:
        nonlocal pos
        kids = []
        if pos < n and s[pos] == "(":
            pos += 1
            while True:
                kids.append(parse())
                if pos >= n:
                    raise ValueError("unmatched braces %s" % s[:100])
                if s[pos] == ",":
                    pos += 1
                elif s[pos] == ")":
                    pos += 1
                    break
                else:
                    raise ValueError("unexpected character %r in tree" % s[pos])
        start = pos
        while pos < n and s[pos] not in ",()":
            pos += 1
        node = _parse_label(s[start:pos])
        for k in kids:
            k.parent = node
        node.children = kids
        return node

    root = parse()
    if pos != n:
        raise ValueError("unmatched braces %s" % s[:100])
    return root
