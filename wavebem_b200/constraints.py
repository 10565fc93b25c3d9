"""Host-side restatement of BEMProblem<3>::compute_constraints
(reference source/bem_problem.cc:990-1105), producing the flattened ConstraintMatrix lines
that the C ABI takes through wbem_set_constraints().

In the drop-in this stays the reference's own host code (it needs deal.II's
ConstraintMatrix, compute_normals and compute_surface_gradients); here it exists so that the
synthetic benchmark/test meshes carry the same kind of constraint lines (double-node
equalities and Dirichlet-value inhomogeneities).  It is O(N) sparse work, not a kernel.
"""
from __future__ import annotations

import numpy as np


class ConstraintLines:
    """lines[k]: constrained dof; entries of line k: col[ptr[k]:ptr[k+1]], val[...];
    inhom[k]: inhomogeneity  (x_line = sum val*x_col + inhom)."""

    def __init__(self, n, lines, ptr, col, val, inhom):
        self.n = int(n)
        self.lines = np.ascontiguousarray(lines, dtype=np.uint32)
        self.ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.inhom = np.ascontiguousarray(inhom, dtype=np.float64)

    @property
    def n_lines(self):
        return len(self.lines)

    @staticmethod
    def empty(n):
        return ConstraintLines(n, [], [0], [], [], [])


def resolve_chains(lines):
    """ConstraintMatrix::close() (bem_problem.cc:1104): an entry whose column is itself a constrained dof
    is replaced by that dof's line (entries and inhomogeneity), until no line refers to a constrained
    dof -- e.g. a flat double d1 -> i of a dof i that a sharp edge turns into an inhomogeneous line, or a
    hanging node whose master is a double node.  Lines that needed no substitution are left as they are;
    substituted lines get their entries merged and sorted by column."""
    for _ in range(16):
        changed = False
        for d, (ent, ih) in list(lines.items()):
            if not any(c in lines and c != d for c, _ in ent):
                continue
            acc = {}
            for c, v in ent:
                if c in lines and c != d:
                    sub_ent, sub_ih = lines[c]
                    for c2, v2 in sub_ent:
                        acc[c2] = acc.get(c2, 0.0) + v * v2
                    ih += v * sub_ih
                else:
                    acc[c] = acc.get(c, 0.0) + v
            lines[d] = [sorted(acc.items()), ih]
            changed = True
        if not changed:
            return
    raise ValueError("constraint lines refer to each other in a cycle")


def compute_constraints(dn_ptr, dn_idx, surface_nodes, tmp_rhs, nodes_normals=None,
                        node_surface_gradients=None, hanging=None):
    """Lines in the order deal.II's ConstraintMatrix would hold them after close() (sorted by
    constrained dof).  `hanging`: optional list of (dof, [(master, weight), ...]) standing in
    for DoFTools::make_hanging_node_constraints (bem_problem.cc:1000).

    Branches (bem_problem.cc:1002-1101), for i the first Dirichlet member of its double-node
    set (or the smallest member when none is Dirichlet):
      * i Dirichlet, double d Dirichlet, equal normals  -> sol_d = sol_i
      * i Dirichlet, double d Dirichlet, other normals  -> both normal derivatives fixed from
        the surface gradients of the two faces (inhomogeneous lines)
      * i Dirichlet, double d Neumann                   -> phi_d = tmp_rhs_i
      * i Neumann (then all doubles are Neumann)        -> phi_d = phi_i
    """
    n = len(surface_nodes)
    lines = {}

    def add_line(d):
        lines.setdefault(int(d), [[], 0.0])

    if hanging:
        for dof, entries in hanging:
            add_line(dof)
            lines[int(dof)][0] = [(int(m), float(w)) for m, w in entries]

    for i in range(n):
        doubles = [int(j) for j in dn_idx[dn_ptr[i]:dn_ptr[i + 1]]]
        first = doubles[0]
        for j in doubles:
            if surface_nodes[j] == 1:
                first = j
                break
        if i != first:
            continue
        others = [j for j in doubles if j != i]
        if surface_nodes[i] == 1:
            for d in others:
                if surface_nodes[d] == 1:
                    if nodes_normals is None:
                        raise ValueError("Dirichlet-Dirichlet double nodes need nodes_normals")
                    ni, nd = nodes_normals[i], nodes_normals[d]
                    if np.linalg.norm(nd - ni) < 1e-4:
                        add_line(d)
                        lines[d][0].append((i, 1.0))
                    else:
                        if node_surface_gradients is None:
                            raise ValueError("edge between Dirichlet faces needs surface gradients")
                        gi, gd = node_surface_gradients[i], node_surface_gradients[d]
                        c = float(ni @ nd)
                        this_ng = (1.0 / (1.0 - c ** 2)) * (gd @ ni + (gi @ nd) * c)
                        other_ng = (1.0 / (1.0 - c ** 2)) * (gi @ nd + (gd @ ni) * c)
                        add_line(i)
                        lines[i][1] = float(this_ng)
                        add_line(d)
                        lines[d][1] = float(other_ng)
                else:
                    add_line(d)
                    lines[d][1] = float(tmp_rhs[i])
        else:
            for d in others:
                add_line(d)
                lines[d][0].append((i, 1.0))

    resolve_chains(lines)
    keys = sorted(lines)
    ptr = [0]
    col, val, inhom = [], [], []
    for k in keys:
        ent, ih = lines[k]
        for c, v in ent:
            col.append(c)
            val.append(v)
        ptr.append(len(col))
        inhom.append(ih)
    return ConstraintLines(n, keys, ptr, col, val, inhom)
