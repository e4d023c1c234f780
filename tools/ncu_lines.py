"""Stall samples of one kernel of an .ncu-rep (ncu --set full --import-source on) summed per SOURCE LINE: the SASS page of the
report gives samples per instruction, nvdisasm --print-line-info of the same build's cubin gives the line of every instruction
(matched by position inside the function).   python tools/ncu_lines.py report.ncu-rep file.cubin <kernel substring> [top]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the disassembly: function sections start with ".text.<mangled>"; instruction lines look like "/*0010*/ ..."
lines, cur_line, cur_file, on = [], None, None, False
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        on = kname in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur_file, cur_line, l.strip()))
print("instructions: report %d, cubin %d" % (len(data), len(lines)))
n = min(len(data), len(lines))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for i in range(n):
    r = data[i]
    s = int(r[ci["# Samples"]]) if r[ci["# Samples"]].isdigit() else 0
    tot += s
    key = (lines[i][0], lines[i][1])
    agg[key][0] += s
    agg[key][1] += int(r[ci["Instructions Executed"]]) if r[ci["Instructions Executed"]].isdigit() else 0
    for h in stall_cols:
        v = r[ci[h]]
        if v.isdigit() and int(v):
            agg[key][2][h[6:]] += int(v)
print("total samples", tot)
for (f, ln), (s, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %5s  %6d samples %5.1f%%  %10d inst  %s" % (f, ln, s, 100.0 * s / max(tot, 1), ex, dict(st.most_common(3))))
