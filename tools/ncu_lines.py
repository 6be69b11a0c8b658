"""Developer tool: correlate an ncu SASS source page (CSV) with nvdisasm -g line info -> samples per CUDA source line.
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > sass.csv ; nvdisasm -g -c file.cubin > dis.txt
   python tools/ncu_lines.py sass.csv dis.txt <mangled-kernel-substring> [topN]"""
import csv, re, sys
sass, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
by_exec = len(sys.argv) > 5 and sys.argv[5] == "exec"      # rank by executed warp instructions instead of samples
addr2line = {}
cur = None; infunc = False
for ln in open(dis):
    if ln.startswith(".text."):
        infunc = kern in ln
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass)))
h = next(r for r in rows if "# Samples" in r)
ia, isamp = h.index("Address"), (h.index("Instructions Executed") if by_exec else h.index("# Samples"))
base = None; per = {}; tot = 0
for r in rows[rows.index(h) + 1:]:
    try: a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]); s = int(r[isamp])
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    per[key] = per.get(key, 0) + s; tot += s
src = {}
for (f, l), s in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    try:
        lines = src.setdefault(f, open(f"/root/repo/multimodal_biometric_fingerprints_palms_b200/csrc/{f}").read().splitlines())
        text = lines[l - 1].strip()[:100]
    except Exception: text = ""
    print(f"{100*s/max(tot,1):5.1f}%  {f}:{l:<4d} {text}")
print("samples:", tot)
