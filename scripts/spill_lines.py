"""usage: spill_lines.py OBJECT KERNEL_SUBSTRING -- local-memory (spill) loads / stores of one kernel by source line"""
import re, subprocess, sys, tempfile, os, glob
from collections import Counter
obj, key = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
out = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout.splitlines()
start = end = None
for i, l in enumerate(out):
    if l.startswith(".text."):
        if start is not None and end is None: end = i
        if key in l and start is None: start = i
end = end or len(out)
cur, cnt = None, Counter()
for l in out[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)))
    if re.search(r"\b(STL|LDL)", l): cnt[cur] += 1
for k, v in sorted(cnt.items(), key=lambda x: x[0] or ("", 0)): print(k, v)
