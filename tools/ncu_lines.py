"""Development aid (no GPU needed): per-source-line warp-stall samples of one kernel from an ncu report.
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> [launch index]
Joins `ncu --page source --csv` (SASS rows) with `nvdisasm -g` line info of libcml_b200/libcmlba.so."""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
mangled = None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "libcml_b200", "libcmlba.so")], cwd=tmp, capture_output=True)
sass = []          # one cubin per translation unit: search all of them
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    sass += subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
short = re.sub(r"\(.*", "", kname).split("::")[-1].replace("void ", "").split("<")[0]
# pick the section whose name contains the short kernel name (first template instance = <false>)
start = [i for i, l in enumerate(sass) if l.startswith("//-") and ".text." in l and short in l][0]
off2line, line = {}, None
for l in sass[start + 1:]:
    if l.startswith("//-") and ".text." in l: break
    m = re.search(r'## File "[^"]+", line (\d+)', l)
    if m: line = int(m.group(1)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m: off2line[int(m.group(1), 16)] = line
H = rows[1]; ai, si, ii = H.index("Address"), H.index("# Samples"), H.index("Instructions Executed")
per = collections.defaultdict(lambda: [0, 0]); base = None; tot = toti = 0
for r in rows[2:]:
    if len(r) <= ii or not r[ai].startswith("0x"): continue
    a = int(r[ai], 16); base = a if base is None else base
    ln = off2line.get(a - base); s_, i_ = int(r[si] or 0), int(r[ii] or 0)
    per[ln][0] += s_; per[ln][1] += i_; tot += s_; toti += i_
src = open(os.path.join(ROOT, "libcml_b200", "csrc", "kernels.cuh")).read().split("\n")
print(f"{kname[:80]}: {tot} samples, {toti} warp instructions")
for ln in sorted(k for k in per if k):
    s_, i_ = per[ln]
    if s_ / max(tot, 1) > 0.012 or i_ / max(toti, 1) > 0.02:
        print(f"{ln:5d} {100 * s_ / tot:5.1f}% smp {100 * i_ / toti:5.1f}% ins | {src[ln - 1].strip()[:130]}")
