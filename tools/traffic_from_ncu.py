"""profiles/traffic.json from an `ncu --set full` capture, tied to the build it was taken from.

    ncu -i gpurun_out/r02_step_full.ncu-rep --page raw --csv > profiles/r02_step_full_raw.csv
    python tools/traffic_from_ncu.py profiles/r02_step_full_raw.csv [--build-hash HASH]

Reads dram__bytes_read.sum / dram__bytes_write.sum of every captured launch, averages them per kernel family (headline
flat-gradient step `step_kernel<1, ...>`, run-table step `step_table_kernel<1, ...>`, ...) and records the hash of the
sources the library was built from (bayesdll_b200.build.source_hash(); pass --build-hash when the capture comes from
another checkout).  bench.py reports `roofline.traffic` only while that hash matches the loaded library's.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    from bayesdll_b200 import build
    build_hash = build.source_hash()
    if "--build-hash" in sys.argv:
        build_hash = sys.argv[sys.argv.index("--build-hash") + 1]
        args.remove(build_hash)
    path = args[0]
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols, units = rows[hdr], rows[hdr + 1]
    ix = {c: i for i, c in enumerate(cols)}
    fam = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(cols):
            continue
        name = r[ix["Kernel Name"]]
        rd = float(r[ix["dram__bytes_read.sum"]]) * UNIT[units[ix["dram__bytes_read.sum"]]]
        wr = float(r[ix["dram__bytes_write.sum"]]) * UNIT[units[ix["dram__bytes_write.sum"]]]
        inst = float(r[ix["smsp__inst_executed.sum"]]) if "smsp__inst_executed.sum" in ix else None
        dur = float(r[ix["gpu__time_duration.sum"]]) if "gpu__time_duration.sum" in ix else None
        fam.setdefault(name, []).append((rd, wr, inst, dur))
    out = {"source": os.path.relpath(path, ROOT), "build_hash": build_hash, "kernels": {}}
    for name, v in fam.items():
        n = len(v)
        out["kernels"][name] = {"launches": n, "dram_bytes_read": sum(x[0] for x in v) / n, "dram_bytes_write": sum(x[1] for x in v) / n,
                                "dram_bytes": sum(x[0] + x[1] for x in v) / n,
                                "warp_instructions": None if v[0][2] is None else sum(x[2] for x in v) / n,
                                "duration": None if v[0][3] is None else sum(x[3] for x in v) / n}
    head = [k for k in out["kernels"] if re.search(r"step_kernel<\(?(int\))?1, .*(true|1)>\(", k) and "table" not in k]
    if head:
        k = out["kernels"][head[0]]
        out["sghmc_step_kernel"] = head[0]
        out["sghmc_step_dram_bytes_read_per_launch"] = k["dram_bytes_read"]
        out["sghmc_step_dram_bytes_write_per_launch"] = k["dram_bytes_write"]
        out["sghmc_step_dram_bytes_per_launch"] = k["dram_bytes"]
        out["sghmc_step_algorithmic_bytes_per_launch"] = 24 * 305548325
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}, indent=1))
    for k, v in out["kernels"].items():
        print(f"{v['launches']:3d} x {k[:110]}: {v['dram_bytes'] / 1e9:.3f} GB, {v['warp_instructions']}, {v['duration']}")


if __name__ == "__main__":
    main()
