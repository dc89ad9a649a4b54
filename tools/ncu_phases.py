"""Split an ncu --page source --csv SASS dump of k_train at BAR.SYNC instructions and print samples per segment."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seg = []
cur = dict(n=0, samples=0, inst=0, ffma=0, lds=0, ldg=0, mufu=0, stall={s: 0 for s in stalls}, first=None, excess=0, wav=0)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]]
    s = int(r[ix["# Samples"]] or 0)
    ie = int(r[ix["Instructions Executed"]] or 0)
    cur["n"] += 1
    cur["samples"] += s
    cur["inst"] += ie
    op = src.split()[0] if not src.strip().startswith("@") else src.split()[1]
    if op.startswith("FFMA") or op.startswith("FMUL") or op.startswith("FADD"):
        cur["ffma"] += ie
    if op.startswith("LDS") or op.startswith("STS"):
        cur["lds"] += ie
        cur["excess"] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        cur["wav"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    if op.startswith("LDG") or op.startswith("STG"):
        cur["ldg"] += ie
    if op.startswith("MUFU"):
        cur["mufu"] += ie
    for st in stalls:
        cur["stall"][st] += int(r[ix[st]] or 0)
    if cur["first"] is None:
        cur["first"] = src.strip()[:40]
    if "BAR.SYNC" in src:
        seg.append(cur)
        cur = dict(n=0, samples=0, inst=0, ffma=0, lds=0, ldg=0, mufu=0, stall={s: 0 for s in stalls}, first=None, excess=0, wav=0)
seg.append(cur)
tot = sum(s["samples"] for s in seg)
print("total samples", tot, "segments", len(seg))
print("%3s %6s %6s %9s %9s %8s %8s %7s %8s %8s  top stalls" % ("seg", "sass", "smp%", "inst", "ffma", "lds", "ldg", "mufu", "smwav", "smexc"))
for i, s in enumerate(seg):
    top = sorted(s["stall"].items(), key=lambda kv: -kv[1])[:3]
    print("%3d %6d %6.2f %9d %9d %8d %8d %7d %8d %8d  %s" % (i, s["n"], 100.0 * s["samples"] / max(tot, 1), s["inst"], s["ffma"], s["lds"], s["ldg"], s["mufu"], s["wav"], s["excess"],
          " ".join("%s=%d" % (k[6:], v) for k, v in top)))
