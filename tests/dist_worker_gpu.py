"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run, one rank per GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wavebem_b200 as wb  # noqa: E402
from wavebem_b200 import dist as wd  # noqa: E402
from wavebem_b200 import meshgen  # noqa: E402
from wavebem_b200.constraints import compute_constraints  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank, world, local = wd.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    m = meshgen.wigley_tank(nxm=14, nt=6, nxu=5, nxd=7, nz=3, nzh=4)
    bc = meshgen.towing_tank_bc(m)
    nn = meshgen.cell_normals_at_nodes(m)
    cl = compute_constraints(m.dn_ptr, m.dn_idx, m.surface_nodes, bc, nodes_normals=nn)
    kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ctx = wb.Context(device=local, rank=rank, world_size=world, gmres_tol=1e-12, gmres_max_steps=400,
                     precond_kind=kind)
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    assert (ctx.row0, ctx.row1) == wd.row_block(m.n_nodes, rank, world)
    wd.init_comm(ctx)
    use_p2p = len(sys.argv) > 2 and sys.argv[2] == "p2p"
    if use_p2p:
        assert wd.init_peer_gather(ctx), "CUDA IPC peer gather unavailable"
    ctx.set_geometry(m.xyz)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(cl)
    rows_n = ctx.get_rows(0)
    alpha = ctx.get_alpha()
    x = np.sin(0.37 * np.arange(m.n_nodes))
    y = ctx.constrained_vmult(x)
    z = np.zeros(m.n_nodes)
    phi, dphi, it, res = ctx.solve_system(z, z, bc)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), rows_n=rows_n, alpha=alpha, y=y, phi=phi, dphi=dphi,
             it=it, res=res, block=np.array([ctx.row0, ctx.row1]))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
