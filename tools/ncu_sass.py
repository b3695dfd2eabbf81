"""print SASS rows (with executed count and top stall reasons) for given source lines of a kernel in an ncu report"""
import csv, os, re, subprocess, sys, tempfile
rep, so, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines = []; cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for k, r in enumerate(rows[2:]):
    f, ln = lines[k] if k < len(lines) else ("?", 0)
    if f == fname and lo <= ln <= hi:
        st = sorted(((int(float(r[ix[c]] or 0)), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{ln:4d} {int(float(r[ix['Instructions Executed']] or 0)):9d} smp {int(float(r[ix['# Samples']] or 0)):6d} {r[ix['Source']][:70]:70s} {[s for s in st if s[0]]}")
