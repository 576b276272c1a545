"""Group an ncu gpu__time_duration CSV (one c3 step) by kernel: launches, total ns, share."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
scale = {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("<unnamed>::", "").strip()
    tot[name] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1); cnt[name] += 1
T = sum(tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("kernel,launches,total_ns,share")
for k, v in tot.most_common():
    print(f'"{k}",{cnt[k]},{int(v)},{v / T * 100:.2f}%')
print(f"TOTAL,{sum(cnt.values())},{int(T)},100%")
