"""Aggregates the per-instruction warp-stall samples of an ncu report (source page) by code region and opcode.
usage: python scripts/ncu_source_hotspots.py <rep> [--top]   (build box, no GPU)"""
import collections, csv, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ins = []
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16)
    base = a if base is None else base
    ins.append((a - base, r[col["Source"]].strip(), int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0),
                {s: int(r[col[s]] or 0) for s in stalls}))
total = sum(x[2] for x in ins)
# regions: split at out-of-line callees (targets of CALL)
targets = sorted({int(m.group(1), 16) - base for _, s, *_ in ins for m in [re.search(r"CALL\.\S+ (0x[0-9a-f]+)", s)] if m})
bounds = [0] + targets + [ins[-1][0] + 16]
print(f"total samples {total}; callees at {[hex(t) for t in targets]}")
for k in range(len(bounds) - 1):
    seg = [x for x in ins if bounds[k] <= x[0] < bounds[k + 1]]
    smp = sum(x[2] for x in seg)
    ex = sum(x[3] for x in seg)
    st = collections.Counter()
    for x in seg:
        st.update(x[4])
    top = ", ".join(f"{s[6:]} {v * 100 / max(smp, 1):.0f}%" for s, v in st.most_common(6))
    print(f"region {k} [{bounds[k]:#x},{bounds[k+1]:#x}): samples {smp * 100 / total:.1f}%  executed {ex}  samples/inst {smp / max(ex, 1) * 1e3:.2f}e-3  | {top}")
    byop = collections.Counter()
    exop = collections.Counter()
    for x in seg:
        op = re.sub(r"^@!?U?P\d\s+", "", x[1]).split()[0] if x[1] else "?"
        byop[op] += x[2]
        exop[op] += x[3]
    print("    by opcode: " + ", ".join(f"{o} {v * 100 / total:.1f}% ({exop[o]})" for o, v in byop.most_common(10)))
if "--top" in sys.argv:
    for x in sorted(ins, key=lambda x: -x[2])[:40]:
        st = ", ".join(f"{s[6:]} {v}" for s, v in sorted(x[4].items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"  {x[0]:05x} {x[2]:6d} {x[1][:60]:60s} {st}")
