"""Join an ncu SASS source page (csv) with nvdisasm -g line info to aggregate executed instructions and
stall samples per CUDA source line.  usage: ncu_lines.py <ncu-rep> <lib.so> <kernel-substr> [top]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines = []   # per instruction: (file, line)
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
print("instructions: nvdisasm", len(lines), "ncu", len(body))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for k, r in enumerate(body):
    key = lines[k] if k < len(lines) else ("?", 0)
    ie = int(float(r[ix["Instructions Executed"]] or 0)); sm = int(float(r[ix["# Samples"]] or 0))
    agg[key][0] += ie; agg[key][1] += sm; agg[key][2] += 1
    tot[0] += ie; tot[1] += sm
print("total warp-instr", tot[0], "samples", tot[1])
srcs = {}
def src(f, n):
    if f not in srcs:
        p = [os.path.join(d, f) for d in ("trajectory_generation_b200/csrc", "include") if os.path.exists(os.path.join(d, f))]
        srcs[f] = open(p[0]).read().splitlines() if p else []
    return srcs[f][n - 1].strip()[:90] if 0 < n <= len(srcs[f]) else ""
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0]}:{key[1]:<5d} instr {100*v[0]/tot[0]:5.1f}%  samples {100*v[1]/tot[1]:5.1f}%  sass {v[2]:5d} | {src(*key)}")
# per file region summary
reg = collections.defaultdict(lambda: [0, 0])
for key, v in agg.items():
    reg[key[0]][0] += v[0]; reg[key[0]][1] += v[1]
for f, v in reg.items():
    print(f, f"instr {100*v[0]/tot[0]:.1f}% samples {100*v[1]/tot[1]:.1f}%")
