"""Per-source-line summary (stall samples, warp instructions) of one kernel in an .ncu-rep.
usage: ncu_lines.py report.ncu-rep kernel-regex[:instance] [topN]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-id", f"::regex:{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur_file = ""
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        lines.append((cur_file, int(r[0]), r[1].strip(), int(d["# Samples"]), int(d["Instructions Executed"]), d))
    except ValueError:
        pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: 0 for k in stall_keys}
for l in lines:
    for k in stall_keys:
        try: agg[k] += int(l[5][k])
        except ValueError: pass
print("stalls:", ", ".join(f"{k[6:]}={v*100//max(1,sum(agg.values()))}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, n, src, s, i, d in sorted(lines, key=lambda l: -l[3])[:top]:
    st = sorted(((int(d[k]) if d[k].isdigit() else 0, k[6:]) for k in stall_keys), reverse=True)[:3]
    print(f"{s*100/tot_s:5.1f}% smp {i*100/tot_i:5.1f}% ins  {f}:{n:<4d} {src[:70]:70s} {' '.join(f'{k}:{v}' for v,k in st if v)}")
