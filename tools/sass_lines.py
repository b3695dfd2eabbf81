"""print the SASS (memory/FP64 instructions only by default) attributed to a source line range of a kernel"""
import os, re, subprocess, sys, tempfile
so, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
pat = re.compile(sys.argv[6]) if len(sys.argv) > 6 else re.compile(r"LDS|STS|DFMA|DMUL|DADD|BAR|SHFL|MUFU|CALL|BRA")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur[0] == fname and lo <= cur[1] <= hi and pat.search(m.group(2)):
        print(cur[1], m.group(1), m.group(2).strip()[:70])
