"""Top stall locations of one kernel from an .ncu-rep (SASS view).
    python scripts/ncu_hot.py rep.ncu-rep <kernel regex> [nlines]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel instance only
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
        continue
    data.append(r)
si, ni, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[ni]) for r in data)
texec = sum(int(r[ei]) for r in data)
print("instructions", len(data), "samples", tot, "warp-instr executed", texec)
cnt, cex = {}, {}
for r in data:
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    op = op.split(".")[0]
    cnt[op] = cnt.get(op, 0) + int(r[ni])
    cex[op] = cex.get(op, 0) + int(r[ei])
print("opcode            samples   %   executed  %")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:16]:
    print("%-16s %8d %5.1f %10d %5.1f" % (k, v, 100.0 * v / max(tot, 1), cex[k], 100.0 * cex[k] / max(texec, 1)))
print("--- top stall lines (index, samples, executed, sass)")
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][ni]))[:nl]:
    print(idx, r[ni], r[ei], r[si][:100])
