"""Split `cuobjdump -sass` of libcfd_b200.so per kernel into /tmp/sass/<name>.sass and print instruction histograms."""
import os
import re
import subprocess
import sys
import collections

so = sys.argv[1] if len(sys.argv) > 1 else "compact_finite_differences_b200/libcfd_b200.so"
out = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
os.makedirs("/tmp/sass", exist_ok=True)
cur, bufs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
        cur = re.sub(r"\(.*", "", cur).replace("void cfd::", "").replace(" ", "")
        bufs[cur] = []
    elif cur:
        bufs[cur].append(line)
for name, lines in bufs.items():
    open(f"/tmp/sass/{name}.sass", "w").write("\n".join(lines))
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    print(name, sum(ops.values()), dict(ops.most_common(14)))
