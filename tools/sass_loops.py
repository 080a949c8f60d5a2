#!/usr/bin/env python
"""Static view of one kernel's SASS: instruction count, loops (backward branches) with their size, local-memory
accesses and MUFU counts.  CPU only (cuobjdump / nvdisasm).
    python tools/sass_loops.py k_visibilityENS [path/to/libocc_b200.so]"""
import os
import re
import subprocess
import sys
import tempfile

name = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "objectcentricocccompletion_b200", "csrc", "libocc_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "annotate", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
text = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(text) if l.strip().startswith(".section") and ".text." in l and name in l]
i0 = starts[0]
i1 = next((i for i in range(i0 + 1, len(text)) if text[i].strip().startswith(".section")), len(text))
body = text[i0:i1]
ins = [(i, l) for i, l in enumerate(body) if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
print("instructions", len(ins), " local ld/st", sum(1 for _, l in ins if "LDL" in l or "STL" in l))
labels = {l.split(":")[0]: i for i, l in enumerate(body) if re.match(r"\.L_x_\d+:", l)}
for i, l in ins:
    m = re.search(r"BRA\S* .*`\((\.L_x_\d+)\)", l)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        seg = [x for j, x in ins if labels[m.group(1)] <= j <= i]
        print("loop %-10s %5d instr  local %3d  MUFU.RSQ %2d  LDS %3d  LDG %3d  FFMA %3d" % (
            m.group(1), len(seg), sum(1 for x in seg if "LDL" in x or "STL" in x), sum(1 for x in seg if "MUFU.RSQ" in x),
            sum(1 for x in seg if "LDS" in x), sum(1 for x in seg if "LDG" in x), sum(1 for x in seg if "FFMA" in x)))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(body))
