"""Row-sharded run on 2 GPUs (NCCL all-gather inside libwbem) against the 1-GPU result."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gather,kind", [("nccl", 0), ("p2p", 0), ("nccl", 1), ("p2p", 1)])
def test_two_gpus_match_one(wb, tmp_path, gather, kind):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from wavebem_b200 import meshgen
    from wavebem_b200.constraints import compute_constraints
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker_gpu.py"), str(tmp_path), gather, str(kind)]
    subprocess.run(cmd, check=True, timeout=600)
    m = meshgen.wigley_tank(nxm=14, nt=6, nxu=5, nxd=7, nz=3, nzh=4)
    bc = meshgen.towing_tank_bc(m)
    nn = meshgen.cell_normals_at_nodes(m)
    cl = compute_constraints(m.dn_ptr, m.dn_idx, m.surface_nodes, bc, nodes_normals=nn)
    ctx = wb.Context(gmres_tol=1e-12, gmres_max_steps=400, precond_kind=kind)
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_geometry(m.xyz)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(cl)
    rows = ctx.get_rows(0)
    y = ctx.constrained_vmult(np.sin(0.37 * np.arange(m.n_nodes)))
    z = np.zeros(m.n_nodes)
    phi, dphi, it, res = ctx.solve_system(z, z, bc)
    for r in range(2):
        d = np.load(tmp_path / f"r{r}.npz")
        r0, r1 = d["block"]
        assert np.array_equal(d["rows_n"], rows[r0:r1])     # same kernel, same rows: bitwise
        assert np.array_equal(d["alpha"], ctx.get_alpha())
        assert np.array_equal(d["y"], y)
        assert int(d["it"]) == it and np.array_equal(d["phi"], phi) and np.array_equal(d["dphi"], dphi)
    ctx.close()
