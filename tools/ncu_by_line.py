"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line info of build/decoder.o and
aggregate executed instructions / stall samples per source line and per device function of decoder.cu.
usage: python tools/ncu_by_line.py src.csv [kernel-substr] [top-n]"""
import csv, os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "gemma.ggml_b200", "build", "decoder.o")
want = sys.argv[2] if len(sys.argv) > 2 else "decoder_kernelILi2ELb0ELb0"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(ROOT, "gemma.ggml_b200", "csrc", "decoder.cu")).read().split("\n")
# function line ranges from the source
funcs = []
for i, l in enumerate(src):
    m = re.match(r"(?:template.*>\s*)?(?:__device__|__global__|static|inline|__forceinline__|\s)*[\w:<>\*& ]+?\b(\w+)\s*\(", l)
    if (l.startswith("__device__") or l.startswith("__global__") or l.startswith("template")) and m:
        pass
for i, l in enumerate(src):
    m = re.search(r"\b(\w+)\s*\([^;]*$", l)
    if m and (l.startswith("__device__") or l.startswith("__global__") or (i > 0 and src[i-1].startswith("template") and not l.startswith(" "))):
        funcs.append((i + 1, m.group(1)))
def func_of(line):
    name = "?"
    for ln, f in funcs:
        if ln <= line: name = f
        else: break
    return name
sec = None; line = None; insts = []
for l in txt:
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
    if m: sec = m.group(1); continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        line = int(m.group(2)) if m.group(1).endswith("decoder.cu") else -1
        continue
    if not (sec and want in sec): continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m: insts.append((int(m.group(1), 16), line, m.group(2)))
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
base = int(body[0][ci["Address"]], 16) if body[0][ci["Address"]].startswith("0x") else int(body[0][ci["Address"]])
assert len(body) == len(insts), (len(body), len(insts))
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()]); by_fn = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = tot_s = 0
for r, (off, ln, sass) in zip(body, insts):
    ie = int(r[ci["Instructions Executed"]] or 0); sm = int(r[ci["# Samples"]] or 0)
    tot_i += ie; tot_s += sm
    for key, tab in ((ln, by_line), (func_of(ln) if ln and ln > 0 else "other", by_fn)):
        tab[key][0] += ie; tab[key][1] += sm
        for h in stall_cols:
            v = int(r[ci[h]] or 0)
            if v: tab[key][2][h[6:]] += v
print("total inst %d samples %d" % (tot_i, tot_s))
print("== by function")
for f, (ie, sm, st) in sorted(by_fn.items(), key=lambda x: -x[1][1]):
    print("%-22s inst %5.1f%% samples %5.1f%%  %s" % (f, 100 * ie / tot_i, 100 * sm / tot_s, dict(st.most_common(4))))
print("== by line (top %d by samples)" % topn)
for ln, (ie, sm, st) in sorted(by_line.items(), key=lambda x: -x[1][1])[:topn]:
    s = src[ln - 1].strip()[:90] if ln and ln > 0 else "(other file)"
    print("%5s inst %4.1f%% smp %4.1f%% %-16s %s | %s" % (ln, 100 * ie / tot_i, 100 * sm / tot_s, func_of(ln) if ln and ln > 0 else "", dict(st.most_common(2)), s))
