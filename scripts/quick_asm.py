"""Quick assembly timing + parity spot check on the GPU box (development helper).
usage: quick_asm.py [nodes] [--variant V] [--check] [--compare]
  --check    64 rows against the oracle
  --compare  row slabs and alpha of the stream kernel (variant 0) against the colour-per-launch kernel (variant 2), bitwise"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wavebem_b200 as wb
from wavebem_b200 import meshgen
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 20000
variant = int(sys.argv[sys.argv.index("--variant") + 1]) if "--variant" in sys.argv else 0
m = meshgen.wigley_tank_for_nodes(n)


def make(variant):
    ctx = wb.Context(assemble_variant=variant)
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_geometry(m.xyz)
    return ctx


ctx = make(variant)
ms = ctx.time_assemble(5)
t = ctx.timings()
evals = 16.0 * m.n_nodes * m.n_cells
print(f"{os.path.basename(wb.LIB_PATH)} variant {variant}: N={m.n_nodes} assemble {ms:.3f} ms regular {t['assemble_regular_ms']:.3f} ms -> "
      f"{34 * evals / (t['assemble_regular_ms'] * 1e-3) / 1e12:.2f} TFLOP/s algorithmic; peak {ctx.measure_fp64_peak():.2f}", flush=True)
if "--check" in sys.argv:
    from oracle import oracle as orc
    r0 = m.n_nodes // 3
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, r0, r0 + 64)
    gn, gd = ctx.get_rows(0, r0, r0 + 64), ctx.get_rows(1, r0, r0 + 64)
    print("   max abs err N", np.abs(gn - on).max(), "D rel", (np.abs(gd - od) / np.abs(od).max()).max(), flush=True)
if "--compare" in sys.argv:
    other = make(2 if variant == 0 else 0)
    other.assemble()
    ctx.assemble()
    worst = 0
    N = m.n_nodes
    for r0 in sorted({0, N // 5, N // 2, max(0, N - 300)}):
        r1 = min(N, r0 + 300)
        for which in (0, 1):
            a, b = ctx.get_rows(which, r0, r1), other.get_rows(which, r0, r1)
            nd = int((a != b).sum())
            worst = max(worst, nd)
            if nd:
                print(f"   rows [{r0},{r1}) matrix {which}: {nd} entries differ, max |diff| {np.abs(a - b).max():.3e}")
    da = np.abs(ctx.get_alpha() - other.get_alpha()).max()
    print(f"   stream vs colour kernel: {'BITWISE EQUAL' if worst == 0 else 'DIFFERENT'} on the slabs; alpha max diff {da:.3e}", flush=True)
