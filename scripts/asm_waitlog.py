"""Development: which work items of the stream kernel wait for their predecessors (needs a -DWBEM_DBG_WAITLOG build)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wavebem_b200 as wb
from wavebem_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = meshgen.wigley_tank_for_nodes(n)
ctx = wb.Context()
ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
ctx.set_geometry(m.xyz)
ctx.assemble(); ctx.assemble()
st = np.zeros(9)
wb.lib().wbem_plan_check(C.c_uint32(m.n_nodes), C.c_uint32(m.n_cells), np.ascontiguousarray(m.cells, dtype=np.uint32).ctypes.data_as(C.c_void_p),
                         C.c_uint32(32), C.c_uint32(64), st.ctypes.data_as(C.c_void_p))
ncl = int(st[0]); tiles = (m.n_nodes + 127) // 128
buf = np.zeros(tiles * ncl, dtype=np.uint32)
wb.lib().wbem_debug_read_asm_sync.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
rc = wb.lib().wbem_debug_read_asm_sync(ctx._h, 4 + tiles * ncl, tiles * ncl, buf.ctypes.data_as(C.c_void_p))
spins = (buf & 0xfffff).reshape(tiles, ncl); cta = (buf >> 20).reshape(tiles, ncl)
print("rc", rc, "items", spins.size, "waited", int((spins > 0).sum()), "total spins", int(spins.sum()), "max", int(spins.max()))
t, k = np.nonzero(spins > 20)
print("items with > 20 spins:", len(t))
print("by tile % G(=env):", np.bincount(t % int(os.environ.get("WBEM_ASM_GROUP", "3")), minlength=1))
print("kpos histogram (10 bins):", np.histogram(k, bins=10, range=(0, ncl))[0])
print("tile histogram (10 bins):", np.histogram(t, bins=10, range=(0, tiles))[0])
order = np.argsort(-spins.ravel().astype(np.int64))[:30]
for o in order:
    print("  tile", o // ncl, "kpos", o % ncl, "spins", spins.ravel()[o], "cta", cta.ravel()[o])
