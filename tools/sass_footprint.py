"""static code footprint of a kernel per source function / line range: counts SASS instructions by source line
(nvdisasm -g) and aggregates by buckets of the source file.  usage: sass_footprint.py <lib.so> <kernel-substr> [bucket_lines=25]"""
import collections, os, re, subprocess, sys, tempfile
so, kern = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
cur = ("?", 0)
cnt = collections.Counter()
per_line = collections.Counter()
tot = 0
for l in dis[start + 1:]:
    if l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[(cur[0], cur[1] // bucket * bucket)] += 1
        per_line[cur] += 1
        tot += 1
print("total instructions", tot)
for (f, b), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{c:6d} {100*c/tot:5.1f}%  {f}:{b}-{b+bucket-1}")
print("top lines")
for (f, ln), c in per_line.most_common(25):
    print(f"{c:6d} {f}:{ln}")
