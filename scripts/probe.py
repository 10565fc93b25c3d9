"""FP64 issue-path micro-benchmarks (development; lib/libwbem_probes.so, csrc/probes.cu)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavebem_b200 import build  # noqa: E402

lib = C.CDLL(build.build_probes())


def probe(n):
    v = C.c_double(0)
    rc = lib.wbem_probe(0, int(n), C.byref(v))
    assert rc == 0, rc
    return v.value


for n, name in ((100, "DFMA"), (101, "DADD"), (102, "DMUL"), (103, "DFMA/DADD alternating")):
    print("opcode probe: 8 chains of %s -> %.2f T(2*instr)/s" % (name, probe(n)))
print("dmma probe: 4 independent m8n8k4 FP64 MMAs -> %.2f TFLOP/s on the tensor pipe" % probe(104))
for n, name in ((107, "8 DFMA"), (105, "8 DFMA + 1 DMMA"), (106, "8 DFMA + 2 DMMA")):
    print("dmma probe: %s per iteration -> DFMA part runs at %.2f TFLOP/s" % (name, probe(n)))
for n in (0, 2, 4, 8, 16):
    print("issue probe: %2d int instr per 8 DFMA -> %.2f TFLOP/s" % (2 * n, probe(n)))
