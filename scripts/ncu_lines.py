"""Per-source-line summary (stall samples, warp instructions) of one kernel in an .ncu-rep.
usage: ncu_lines.py report.ncu-rep kernel-regex [instance] [topN]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split into kernel instances: a new instance starts at each "Function Name" row that follows data
instances, cur, cur_file, hdr = [], None, "", None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        if cur is None or (cur and cur["seen_main"] and cur_file.endswith(cur["main"])):
            pass
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    instances.append((cur_file, d))
# instances are concatenated per kernel launch; detect restarts by (file,line) repetition
launches, seen, cur = [], set(), []
for f, d in instances:
    key = (f, d["Line No"])
    if key in seen:
        launches.append(cur); cur = []; seen = set()
    seen.add(key); cur.append((f, d))
launches.append(cur)
print(f"{len(launches)} launches matched; showing #{inst}")
lines = []
for f, d in launches[inst]:
    try:
        lines.append((f, int(d["Line No"]), d["Source"].strip(), int(d["# Samples"]), int(d["Instructions Executed"]), d))
    except ValueError:
        pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
stall_keys = [k for k in hdr if k.startswith("stall_")]
agg = {k: 0 for k in stall_keys}
for l in lines:
    for k in stall_keys:
        try: agg[k] += int(l[5][k])
        except (ValueError, KeyError): pass
print("stalls:", ", ".join(f"{k[6:]}={v*100//max(1,sum(agg.values()))}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
import os
skey = 3 if os.environ.get("SORT") == "smp" else 4
for f, n, src, s, i, d in sorted(lines, key=lambda l: -l[skey])[:top]:
    st = sorted(((int(d[k]) if d.get(k, "").isdigit() else 0, k[6:]) for k in stall_keys), reverse=True)[:3]
    print(f"{s*100/tot_s:5.1f}% smp {i*100/tot_i:5.1f}% ins  {f}:{n:<4d} {src[:64]:64s} {' '.join(f'{k}:{v}' for v,k in st if v)}")
# ---- phase totals (file, first line, last line, label) given as extra args "file:lo-hi=label"
if len(sys.argv) > 5:
    print("phases:")
    rest = tot_i
    for spec in sys.argv[5:]:
        rng, label = spec.split("=")
        f, lh = rng.split(":")
        lo, hi = (int(v) for v in lh.split("-"))
        ins = sum(l[4] for l in lines if l[0] == f and lo <= l[1] <= hi)
        smp = sum(l[3] for l in lines if l[0] == f and lo <= l[1] <= hi)
        rest -= ins
        print(f"  {label:28s} {ins*100/tot_i:5.1f}% ins ({ins/1e6:7.2f} M)  {smp*100/tot_s:5.1f}% smp")
    print(f"  {'(other)':28s} {rest*100/tot_i:5.1f}% ins")
