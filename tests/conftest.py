import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/wbem_oracle.c) -- the checker, never the product."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def wb():
    import wavebem_b200
    if not os.path.exists(wavebem_b200.LIB_PATH):
        from wavebem_b200 import build
        build.build()
    return wavebem_b200


def rel_err_rowscaled(a, b, diag=None):
    """Parity metric for matrix entries (DESIGN.md "Parity metric"):
        |a_ij - b_ij| / max(|a_ij|, |b_ij|, row_scale_i),   row_scale_i = max(max_j |b_ij|, |diag_i|).
    Double-layer entries between coplanar panels are analytically 0, and the integrand
    (R.n)/r^3 of near-coplanar neighbours carries an ABSOLUTE rounding floor ~1e-16..1e-14 in the
    reference arithmetic itself, so a purely relative error is meaningless there.  For the
    Neumann matrix the natural row scale is the operator row N + diag(alpha) that enters the
    solve (pass diag=alpha, alpha_i >= 1/8); for the Dirichlet matrix it is max_j |D_ij|."""
    a = np.asarray(a)
    b = np.asarray(b)
    row = np.abs(b).max(axis=-1, keepdims=True)
    if diag is not None:
        row = np.maximum(row, np.abs(np.asarray(diag)).reshape(row.shape))
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), row)
    scale = np.where(scale == 0, 1.0, scale)
    return float((np.abs(a - b) / scale).max())


def make_problem(mesh, orc_mod=None, bc=None):
    """Boundary data + constraint lines for a mesh (host logic shared by oracle and GPU)."""
    from wavebem_b200 import meshgen
    from wavebem_b200.constraints import compute_constraints
    if bc is None:
        bc = meshgen.towing_tank_bc(mesh)
    nn = meshgen.cell_normals_at_nodes(mesh)
    cl = compute_constraints(mesh.dn_ptr, mesh.dn_idx, mesh.surface_nodes, bc, nodes_normals=nn)
    return bc, nn, cl
