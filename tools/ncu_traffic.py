"""Not a test: turn an ncu launch list captured with NVTX renaming into the per-family tables the
bench and DESIGN.md cite.

    # on the GPU box (the same command first without ncu, then under it):
    python bench.py --ncu-step cfg2 &&
    ncu --nvtx --print-nvtx-rename kernel --profile-from-start off --clock-control none \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --csv --log-file gpurun_out/r02_ncu_cfg2_families.csv python bench.py --ncu-step cfg2
    # here:
    python tools/ncu_traffic.py gpurun_out/r02_ncu_cfg2_families.csv cfg2

Every libmmemo launch of the profiled step runs inside an NVTX range named after bench.py's kernel
family (entry point + shape), so the "Kernel Name" column IS the family.  Writes
profiles/r02_ncu_<workload>_families.txt (launches, mean / total device time, share of the step,
DRAM bytes per launch) and merges the DRAM bytes into profiles/ncu_traffic.json, which bench.py
reads for `roofline.traffic`."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, workload):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ci = {n: i for i, n in enumerate(hdr)}
    per = {}
    for r in rows[start + 1:]:
        if len(r) <= ci["Metric Value"]:
            continue
        key = (r[ci["ID"]], r[ci["Kernel Name"]])
        per.setdefault(key, {})[r[ci["Metric Name"]]] = (float(r[ci["Metric Value"]].replace(",", "")),
                                                         r[ci["Metric Unit"]])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
             "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    fam = {}
    for (_, name), m in per.items():
        # "family/kernel name" when the launch ran inside an NVTX range (every libmmemo launch does)
        family, _, kern = name.partition("/")
        f = fam.setdefault(family, dict(n=0, us=0.0, rd=0.0, wr=0.0, kernels={}))
        kern = kern.split("(")[0].split("<")[0] if "<unnamed>::" in kern.split("<")[0] + "<" else kern
        kern = kern.replace("void ", "").replace("<unnamed>::", "").split("<")[0]   # instantiations count together
        f["kernels"][kern] = f["kernels"].get(kern, 0) + 1
        f["n"] = max(f["kernels"].values())        # C-ABI calls = launches of the family's main kernel
        v, u = m.get("gpu__time_duration.sum", (0.0, "us"))
        f["us"] += v * scale.get(u, 1.0)
        v, u = m.get("dram__bytes_read.sum", (0.0, "byte"))
        f["rd"] += v * scale.get(u, 1.0)
        v, u = m.get("dram__bytes_write.sum", (0.0, "byte"))
        f["wr"] += v * scale.get(u, 1.0)
    total = sum(f["us"] for f in fam.values())
    out = os.path.join(ROOT, "profiles", f"r02_ncu_{workload}_families.txt")
    with open(out, "w") as fh:
        fh.write(f"# {workload}: one eager step under ncu (cold caches, serialised launches: compare "
                 f"SHARES, not absolute times); source {os.path.basename(path)}\n")
        fh.write(f"# kernel family (NVTX range of the launching C-ABI call; unnamed = torch kernels)   calls   "
                 f"us/call   share   DRAM read MB/call   DRAM write MB/call   kernels in the range\n")
        for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
            kern = ", ".join(k[:48] for k in f["kernels"] if k)
            fh.write(f"{name[:96]:96s} {f['n']:4d} {f['us'] / f['n']:9.1f} {f['us'] / total:7.3f} "
                     f"{f['rd'] / f['n'] / 1e6:9.2f} {f['wr'] / f['n'] / 1e6:9.2f}   {kern}\n")
        fh.write(f"# total device time of the step's kernels: {total:.1f} us\n")
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(tj)) if os.path.isfile(tj) else {"families": {}, "source": {}}
    for name, f in fam.items():
        if name.split(":")[0].endswith(("_bf16", "_f32")) or "grouped" in name:
            data["families"][name] = (f["rd"] + f["wr"]) / f["n"]
    data.setdefault("source", {})[workload] = f"profiles/r02_ncu_{workload}_families.txt"
    json.dump(data, open(tj, "w"), indent=1, sort_keys=True)
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
