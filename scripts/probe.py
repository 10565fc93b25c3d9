import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavebem_b200 as wb
ctx = wb.Context()
print("fp64 peak", ctx.measure_fp64_peak())
for n, name in ((100, "DFMA"), (101, "DADD"), (102, "DMUL"), (103, "DFMA/DADD alternating")):
    print("opcode probe: 8 chains of %s -> %.2f T(2*instr)/s" % (name, ctx.issue_probe(n)))
print("dmma probe: 4 independent m8n8k4 FP64 MMAs -> %.2f TFLOP/s on the tensor pipe" % ctx.issue_probe(104))
for n, name in ((107, "8 DFMA"), (105, "8 DFMA + 1 DMMA"), (106, "8 DFMA + 2 DMMA")):
    print("dmma probe: %s per iteration -> DFMA part runs at %.2f TFLOP/s" % (name, ctx.issue_probe(n)))
for n in (0, 2, 4, 8, 16):
    print("issue probe: %2d int instr per 8 DFMA -> %.2f TFLOP/s" % (2 * n, ctx.issue_probe(n)))
