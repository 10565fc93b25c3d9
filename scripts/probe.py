import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavebem_b200 as wb
ctx = wb.Context()
print("fp64 peak", ctx.measure_fp64_peak())
for n in (0, 2, 4, 8, 16):
    print("issue probe: %2d int instr per 8 DFMA -> %.2f TFLOP/s" % (2 * n, ctx.issue_probe(n)))
