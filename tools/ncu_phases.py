"""Per-phase (barrier to barrier) stall / shared-memory wavefront summary of one kernel from an ncu report:
    python tools/ncu_phases.py gpurun_out/prof.ncu-rep [kernel-substring]
Splits the SASS of the kernel at BAR.SYNC and sums the source-page counters per segment."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
seg, cur = [], []
for r in data:
    if len(r) < len(hdr): continue
    cur.append(r)
    if "BAR.SYNC" in r[ix["Source"]] or "RET." in r[ix["Source"]]:
        seg.append(cur); cur = []
seg.append(cur)
tot = sum(f(r, "# Samples") for r in data if len(r) >= len(hdr))
print("total samples", tot, "segments", len(seg))
keys = ["stall_barrier", "stall_short_sb", "stall_long_sb", "stall_math", "stall_mio", "stall_wait", "stall_not_selected",
        "stall_selected", "stall_dispatch", "stall_lg", "stall_branch_resolving", "stall_no_inst", "stall_sleep", "stall_membar"]
for k, s in enumerate(seg):
    smp = sum(f(r, "# Samples") for r in s)
    if smp < 0.003 * tot: continue
    wf = sum(f(r, "L1 Wavefronts Shared") for r in s); ideal = sum(f(r, "L1 Wavefronts Shared Ideal") for r in s)
    inst = sum(f(r, "Instructions Executed") for r in s)
    st = {key[6:]: sum(f(r, key) for r in s) for key in keys}
    top = sorted(st.items(), key=lambda x: -x[1])[:5]
    first = s[0][ix["Address"]] if s else ""
    print(f"seg {k:3d} n={len(s):5d} samples {100*smp/tot:5.1f}% inst {inst/1e6:7.2f}M wf {wf/1e6:6.1f}M ideal {ideal/1e6:6.1f}M",
          [(a, round(100 * b / max(smp, 1))) for a, b in top])
