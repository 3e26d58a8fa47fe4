"""Joins an ncu SASS source page with the line info of the cubin: executed warp instructions and stall
samples per source line.   python scratch/ncu_lines.py report.ncu-rep lib.so cubin_name kernel_mangled [top]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + mangled + ":"))
insts = []  # (file, line, text, inline chain)
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
assert len(body) == len(insts), (len(body), len(insts))
per = collections.defaultdict(lambda: [0, 0, 0, 0])
tot_i = tot_s = 0
for (loc, txt), r in zip(insts, body):
    ie = int(r[ix["Instructions Executed"]]); sm = int(r[ix["# Samples"]]); te = int(r[ix["Thread Instructions Executed"]])
    key = (loc[0], loc[1])
    per[key][0] += ie; per[key][1] += sm; per[key][2] += te; per[key][3] += 1
    tot_i += ie; tot_s += sm
print(f"total warp instructions {tot_i}, samples {tot_s}, SASS instructions {len(insts)}")
if os.environ.get("DUMP"):
    for (loc, txt), r in zip(insts, body):
        print(f"{loc[0]}:{loc[1]:<5d} {int(r[ix['Instructions Executed']]):>12d} {int(r[ix['# Samples']]):>7d} {r[ix['Avg. Threads Executed']]:>5s}  {txt}")
    sys.exit()
for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:<5d} inst {100 * v[0] / tot_i:5.2f}%  samples {100 * v[1] / tot_s:5.2f}%  lanes {v[2] / max(v[0], 1):5.1f}  sass {v[3]}")
