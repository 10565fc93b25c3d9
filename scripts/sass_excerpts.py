"""Compact SASS excerpts of the two hot kernels for profiles/ (runs without a GPU: cuobjdump on the built objects).

    python scripts/sass_excerpts.py r2      # writes profiles/r2_sass_k_assemble_rows.txt, profiles/r2_sass_k_bem_gemv.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"


def sass(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for b in re.split(r"(?=\s+Function : )", out):
        m = re.search(r"Function : (\S+)", b)
        if m and re.search(pat, m.group(1)):
            return m.group(1), b
    return None, ""


def clean(b):
    lines = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x", l)
        if m:
            lines.append((m.group(1), m.group(2).strip()))
    return lines


def hist(lines):
    return collections.Counter(re.sub(r"^(@!?U?P\d+\s+)", "", t).split()[0] for _, t in lines)


name, b = sass(os.path.join(ROOT, "wavebem_b200/lib/assemble.o"), r"k_assemble_rows")
L = clean(b)
idx = [i for i, (_, t) in enumerate(L) if "MUFU.RSQ64H" in t]
lo, hi = idx[0] - 75, idx[-1] + 95
with open(os.path.join(ROOT, f"profiles/{tag}_sass_k_assemble_rows.txt"), "w") as f:
    f.write(f"# cuobjdump -sass wavebem_b200/lib/assemble.o, {name} (sm_100a), {len(L)} instructions\n")
    f.write("# opcode histogram of the whole kernel:\n")
    for k, v in hist(L).most_common(40):
        f.write(f"#   {k:28s} {v}\n")
    body = L[lo:hi]
    f.write(f"#\n# the per-cell body of the integration loop (two Gauss lines x 2, then the scatter), {len(body)} instructions:\n")
    f.write("#   " + ", ".join(f"{k} {v}" for k, v in hist(body).most_common(16)) + "\n")
    special = sorted({t.split()[0] for _, t in L if t.startswith(("UBLKCP", "SYNCS", "ATOMG", "ATOMS", "REDG", "MEMBAR", "CCTL", "ERRBAR"))})
    f.write("#   TMA / mbarrier / ordering instructions elsewhere in the kernel: " + ", ".join(special) + "\n#\n")
    for a, t in body:
        f.write(f"/*{a}*/ {t}\n")
name, b = sass(os.path.join(ROOT, "wavebem_b200/lib/operator.o"), r"k_bem_gemv\d+GemvArgs$")
L = clean(b)
with open(os.path.join(ROOT, f"profiles/{tag}_sass_k_bem_gemv.txt"), "w") as f:
    f.write(f"# cuobjdump -sass wavebem_b200/lib/operator.o, {name} (sm_100a), {len(L)} instructions\n# opcode histogram:\n")
    for k, v in hist(L).most_common(30):
        f.write(f"#   {k:28s} {v}\n")
    idx = [i for i, (_, t) in enumerate(L) if "LDG.E.NA.128" in t]
    if idx:
        lo = max(0, idx[0] - 30)
        f.write("#\n# first streaming loop (128-bit matrix loads without L1 allocation, FP64 FMAs):\n")
        for a, t in L[lo:min(len(L), lo + 170)]:
            f.write(f"/*{a}*/ {t}\n")
print("SASS excerpts written")
