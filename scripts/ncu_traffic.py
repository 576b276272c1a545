"""profiles/ncu_traffic.json from an ncu CSV of dram__bytes_{read,write}.sum over one c3 step (bench.py --ncu-range):
average DRAM bytes per launch for each ABI kernel class that bench.py's roofline may name."""
import collections, csv, json, sys

CLASSES = {"mclip_dwconv_backward": ("mclip_dws_bwd", "mclip_dwconv_bwd_kernel"), "mclip_dwconv_forward": ("mclip_dws_fwd", "mclip_dwconv_fwd_kernel"),
           "mclip_ew_backward": ("mclip_ew_bwd_kernel",), "mclip_ew_forward": ("mclip_ew_fwd_kernel",), "mclip_gemm_tn": ("mclip_gemm_tn_kernel",),
           "mclip_gemm_wgrad": ("mclip_gemm_wgrad_kernel",)}


def main(path, out, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    launches = collections.defaultdict(set)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows:
        if r is hdr or r[im] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        for cls, pats in CLASSES.items():
            if any(p in r[ik] for p in pats):
                per[cls][r[im]] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
                launches[cls].add(r[iid])
    res = {}
    for cls, d in per.items():
        n = len(launches[cls])
        res[cls] = {"bytes_per_launch": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / n, "dram_read_bytes_per_launch": d["dram__bytes_read.sum"] / n,
                    "dram_write_bytes_per_launch": d["dram__bytes_write.sum"] / n, "launches_captured": n, "note": note}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: (round(v["bytes_per_launch"] / 1e9, 3), v["launches_captured"]) for k, v in res.items()}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
