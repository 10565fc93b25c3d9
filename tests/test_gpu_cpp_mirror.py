"""The C++ host mirror of BEMProblem<3> (wavebem_b200/csrc/bem_problem_b200.h) driven as a
compiled program, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import make_problem
from wavebem_b200 import meshgen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_bem_problem(wb, orc, tmp_path):
    from wavebem_b200 import build
    exe = build.build_cpp_test()
    m = meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3)
    bc, nn, cl = make_problem(m)
    n = m.n_nodes
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        np.array([n, m.n_cells, len(m.dn_idx), cl.n_lines, len(cl.col)], dtype=np.uint32).tofile(f)
        for a, dt in ((m.xyz, np.float64), (m.cells, np.uint32), (m.dir_flag, np.uint8), (m.dn_ptr, np.uint32),
                      (m.dn_idx, np.uint32), (m.surface_nodes, np.float64), (m.other_nodes, np.float64),
                      (bc, np.float64), (cl.lines, np.uint32), (cl.ptr, np.uint32), (cl.col, np.uint32),
                      (cl.val, np.float64), (cl.inhom, np.float64)):
            np.ascontiguousarray(a, dtype=dt).tofile(f)
    import torch
    ndev = torch.cuda.device_count()
    out = subprocess.run([exe, str(fin), str(fout), str(ndev)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    r = np.fromfile(fout, dtype=np.float64)
    phi, dphi, alpha, checks = r[:n], r[n:2 * n], r[2 * n:3 * n], r[3 * n:]
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    con = orc.Constraints(n, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)
    ref = orc.solve_system(on, od, m.surface_nodes, m.other_nodes, bc, con, np.zeros(n), np.zeros(n), tol=1e-12,
                           max_steps=400)
    s = m.surface_nodes == 1
    assert np.abs(alpha - ref["alpha"]).max() < 1e-12
    assert np.linalg.norm(phi[~s] - ref["phi"][~s]) < 1e-9 * np.linalg.norm(ref["phi"])
    assert np.linalg.norm(dphi[s] - ref["dphi_dn"][s]) < 1e-9 * np.linalg.norm(ref["dphi_dn"])
    assert abs(checks[0] - ref["iters"]) <= 3 and checks[1] < 1e-16 and checks[2] < 1e-9 and checks[3] == 5
    # precond_kind = 1 through the C++ mirror: same solution, at most half the iterations
    assert checks[4] < 1e-9 and 0 < checks[5] <= checks[0] / 2
    # auto_constraints = 1: the library's compute_constraints gives the same lines, hence the same solve
    assert checks[6] < 1e-9 and checks[7] == cl.n_lines
    assert checks[8] == 1.0     # FlatDomain::generate_double_nodes_set reproduces the sets
    # one single-threaded BEMProblem driving several row blocks (n_gpus > 1): bitwise the 1-GPU result
    assert checks[9] == 1.0 and checks[10] == (ndev if ndev >= 2 else 3) and checks[11] == (1.0 if ndev >= 2 else 0.0)
