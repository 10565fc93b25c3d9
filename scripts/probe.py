import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavebem_b200 as wb
ctx = wb.Context()
print("fp64 peak", ctx.measure_fp64_peak())
for n, name in ((100, "DFMA"), (101, "DADD"), (102, "DMUL"), (103, "DFMA/DADD alternating")):
    print("opcode probe: 8 chains of %s -> %.2f T(2*instr)/s" % (name, ctx.issue_probe(n)))
for n in (0, 2, 4, 8, 16):
    print("issue probe: %2d int instr per 8 DFMA -> %.2f TFLOP/s" % (2 * n, ctx.issue_probe(n)))
