"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: name, launches, total us (in launch order)."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    elif r[ui] in ("s", "second"): v *= 1e6
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"<.*", "<>", name)
    name = name.replace("void ", "").replace("fpm::", "")
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += v
s = sum(t[1] for t in tot.values())
for k, (n, v) in tot.items():
    print("%-64s %3d  %9.1f us  %5.1f %%" % (k[:64], n, v, 100 * v / s))
print("%-64s %3d  %9.1f us" % ("TOTAL (serialised, cold clocks)", sum(t[0] for t in tot.values()), s))
