"""Per-kernel SASS hot spots of an .ncu-rep captured with --set full --import-source on: the instructions with the most
warp-stall samples, their dominant stall reason, and the sample share per opcode.
   python tools/ncu_hotspots.py gpurun_out/x.ncu-rep profiles/rNN_sass_hotspots.txt [top]"""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out, kernel, hdr, rows, seen = [], None, None, [], set()


def flush():
    if kernel is None or not rows or kernel in seen:
        return
    seen.add(kernel)
    i_src, i_all = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    i_exec = hdr.index("Instructions Executed")
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[i_all] or 0) for r in rows) or 1
    out.append(f"== {kernel}\n   {len(rows)} SASS instructions, {total} stall samples, "
               f"{sum(int(r[i_exec] or 0) for r in rows)} warp instructions executed")
    by_op = collections.Counter()
    for r in rows:
        by_op[r[i_src].split()[0].rstrip(";") if r[i_src].split() else "?"] += int(r[i_all] or 0)
    out.append("   samples by opcode: " + ", ".join(f"{op} {100.0 * n / total:.1f}%" for op, n in by_op.most_common(10)))
    for r in sorted(rows, key=lambda r: -int(r[i_all] or 0))[:top]:
        n = int(r[i_all] or 0)
        why = sorted(((int(r[i] or 0), h[6:]) for i, h in reasons), reverse=True)[:2]
        out.append(f"   {100.0 * n / total:5.1f}%  {r[i_src].strip()[:70]:70s}  " +
                   ", ".join(f"{h} {c}" for c, h in why if c))


for line in csv.reader(io.StringIO(raw)):
    if not line:
        continue
    if line[0] == "Kernel Name":
        flush()
        kernel, hdr, rows = line[1], None, []
    elif line[0] == "Address":
        hdr = line
    elif hdr is not None and len(line) >= len(hdr) - 2:
        rows.append(line)
flush()
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
