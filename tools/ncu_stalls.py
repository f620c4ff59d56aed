"""Kernel-wide warp-stall breakdown from the source page of an ncu report (sampling), offline."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
tot = collections.Counter()
n = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        for h, v in zip(hdr, r):
            if h.startswith("stall_") and "Not Issued" not in h and v.isdigit():
                tot[h] += int(v)
        n += 1
s = sum(tot.values()) or 1
print("%d SASS rows, %d samples" % (n, s))
for k, v in tot.most_common():
    print("  %-26s %6d  %5.1f%%" % (k, v, 100.0 * v / s))
