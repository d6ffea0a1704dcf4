"""List local-memory (spill) instructions of the decode kernel by source line: python tools/sass_spills.py [kernel-substr]"""
import re, subprocess, sys, os, tempfile
obj = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gemma.ggml_b200", "build", "decoder.o")
want = sys.argv[1] if len(sys.argv) > 1 else "decoder_kernelILi2ELb0"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(os.path.dirname(obj), "..", "csrc", "decoder.cu")).read().split("\n")
sec = fn = None; line = None; hits = {}; count = {}
for l in txt:
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
    if m: sec = m.group(1); fn = "kernel"; continue
    m = re.match(r"\s*\$\S+\$\S*?\d+(matvec_consumer|score_phase|pv_phase|chain_phase)I\S*:", l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File ".*decoder.cu", line (\d+)', l)
    if m: line = int(m.group(1)); continue
    if not (sec and want in sec): continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l):
        count[fn] = count.get(fn, 0) + 1
        if " LDL" in l or " STL" in l:
            hits.setdefault((fn, line), [0, 0])[0 if " STL" in l else 1] += 1
print("instructions per function:", count)
for (f, ln), (st, ld) in sorted(hits.items(), key=lambda x: (x[0][0], x[0][1] or 0)):
    print(f"{f:16s} {ln:5d} STL {st} LDL {ld} | {src[ln-1].strip()[:100] if ln else ''}")
