"""world_size-2 checks of the row-sharding host logic on CPU (gloo).  The oracle stands in
for the per-rank compute: each rank produces its block of rows, the plumbing in
wavebem_b200.dist gathers them, and the result must equal the unsharded computation."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cube, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as orc
    from wavebem_b200 import dist as wd
    from wavebem_b200 import meshgen
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = meshgen.cube(n_cube)
    n = m.n_nodes
    r0, r1 = wd.row_block(n, rank, world)
    nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, r0, r1, nthreads=1)
    full_n = wd.allgather_rows(nm, n)
    x = np.sin(0.37 * np.arange(n))
    y = wd.allgather_rows(orc.fullmatrix_vmult(dm, x, nthreads=1), n)
    uid = wd.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), full_n=full_n, y=y, uid=np.frombuffer(uid, dtype=np.uint8),
             block=np.array([r0, r1]))
    dist.destroy_process_group()


def test_row_block_partition():
    from wavebem_b200.dist import row_block
    for n in (1, 7, 150, 20073):
        for world in (1, 2, 3, 4, 8):
            blocks = [row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            assert max(b[1] - b[0] for b in blocks) == (n + world - 1) // world


@pytest.mark.parametrize("n_cube", [3, 4])
def test_sharded_rows_gather_to_the_unsharded_result(tmp_path, n_cube):
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    from wavebem_b200 import meshgen
    orc.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_cube, str(tmp_path)), nprocs=2, join=True)
    m = meshgen.cube(n_cube)
    nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, nthreads=1)
    y = orc.fullmatrix_vmult(dm, np.sin(0.37 * np.arange(m.n_nodes)), nthreads=1)
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["full_n"], nm)          # every rank holds the full gathered result
        assert np.array_equal(z["y"], y)
        assert bytes(z["uid"]) == bytes(range(128))     # unique-id hand-off
