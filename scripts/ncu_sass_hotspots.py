"""Per-opcode and per-region stall-sample summary of one kernel from an .ncu-rep (source page, SASS view).
usage: python scripts/ncu_sass_hotspots.py rep.ncu-rep <kernel regex> [--listing]"""
import csv, io, subprocess, sys, collections

rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break  # first launch only
    if r[0] == "Address":
        break
    body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
S, X = ci["# Samples"], ci["Instructions Executed"]
tot_s = sum(int(r[S]) for r in body)
tot_x = sum(int(r[X]) for r in body)
print(f"instructions {len(body)}  samples {tot_s}  warp-instructions executed {tot_x}")
by = collections.defaultdict(lambda: [0, 0, 0])
for r in body:
    op = r[ci["Source"]].split()
    op = op[1] if op[0].startswith("@") else op[0]
    op = op.split(".")[0]
    by[op][0] += int(r[S]); by[op][1] += int(r[X]); by[op][2] += 1
print(f"{'opcode':12s} {'samples%':>9s} {'exec%':>7s} {'static':>7s}")
for op, (s, x, c) in sorted(by.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:12s} {100*s/tot_s:9.2f} {100*x/tot_x:7.2f} {c:7d}")
if "--listing" in sys.argv:
    # regions of 64 consecutive instructions
    step = 64
    for i in range(0, len(body), step):
        blk = body[i:i + step]
        s = sum(int(r[S]) for r in blk); x = sum(int(r[X]) for r in blk)
        ops = collections.Counter((r[ci["Source"]].split()[1] if r[ci["Source"]].split()[0].startswith("@") else r[ci["Source"]].split()[0]).split(".")[0] for r in blk)
        top = " ".join(f"{k}:{v}" for k, v in ops.most_common(5))
        print(f"[{i:5d}] samples {100*s/tot_s:6.2f}%  exec {100*x/tot_x:6.2f}%  {top}")
