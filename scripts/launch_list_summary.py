"""ncu launch list (gpu__time_duration.sum per launch) -> per-kernel shares under profiles/.

  python scripts/launch_list_summary.py gpurun_out/x_launches.csv profiles/out.csv "<command>"
"""
import collections
import csv
import io
import sys

src, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
txt = open(src).read()
rd = list(csv.reader(io.StringIO("\n".join(l for l in txt.splitlines() if l.startswith('"')))))
hdr = rd[0]
ik, iv, im, iu = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
agg = collections.OrderedDict()
for r in rd[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    name = r[ik].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(out, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, %s "
            "(cold-cache, serialised: compare shares)\n" % what)
    f.write("kernel,launches,total_ms,share\n")
    for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('"%s",%d,%.3f,%.4f\n' % (k, c, ms, ms / tot))
print(open(out).read())
