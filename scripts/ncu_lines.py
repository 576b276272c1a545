"""Stall samples / executed instructions of an .ncu-rep kernel aggregated per CUDA source line.
usage: python scripts/ncu_lines.py report.ncu-rep object.o mangled_kernel_substring [top]
(ncu's source page lists SASS in program order; nvdisasm -g gives the line of every SASS instruction of the same cubin.)"""
import collections, csv, io, os, re, subprocess, sys, tempfile


def main(rep, obj, kernel, top=40):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    lines, cur, on = [], None, False
    for ln in dis:
        if ln.startswith("//---") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    ist, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    assert len(data) == len(lines), (len(data), len(lines))
    f = lambda v: float(v) if v else 0.0
    st, ex = collections.Counter(), collections.Counter()
    for r, l in zip(data, lines):
        st[l] += f(r[ist]); ex[l] += f(r[iex])
    ts, te = sum(st.values()) or 1, sum(ex.values()) or 1
    src = {}
    for (fn, n), v in st.most_common(top):
        if fn not in src:
            for root in ("mammo-clip_b200/csrc", "."):
                p = os.path.join(root, fn)
                if os.path.exists(p):
                    src[fn] = open(p).read().splitlines()
                    break
        text = src.get(fn, [""] * (n + 1))[n - 1].strip()[:100] if fn in src else ""
        print(f"{v / ts * 100:5.1f}% stall {ex[(fn, n)] / te * 100:5.1f}% exec  {fn}:{n:<4d} {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
