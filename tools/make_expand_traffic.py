#!/usr/bin/env python3
"""Regenerate profiles/expand_traffic.json from the `ncu --set full --page raw --csv` exports of ONE-witness launches of
the two shipped expand kernels (tools/gpu_r02a.sh).  bench.py reads the JSON for `roofline.traffic`.
    python tools/make_expand_traffic.py profiles/r02a_k_expand_round_ncu_raw.csv profiles/r02a_k_expand_codes_ncu_raw.csv"""
import csv, json, os, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}


def grab(fn):
    hdr, units, vals = list(csv.reader(open(fn)))[:3]
    col = {h: i for i, h in enumerate(hdr)}
    g = lambda k: float(vals[col[k]]) * UNIT[units[col[k]]]
    return {"kernel": vals[col["Kernel Name"]], "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
            "duration_s": g("gpu__time_duration.sum"), "registers": int(float(vals[col["launch__registers_per_thread"]])),
            "dram_write_pct_of_peak": float(vals[col["dram__bytes_write.sum.pct_of_peak_sustained_elapsed"]]),
            "dram_read_pct_of_peak": float(vals[col["dram__bytes_read.sum.pct_of_peak_sustained_elapsed"]])}


def main(argv):
    ks = [grab(f) for f in argv]
    total = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in ks)
    algo = 32 * 215907954
    out = {"dram_bytes_per_witness": total, "algorithmic_bytes_per_witness": algo, "traffic_over_algorithmic": total / algo,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of the two kernels that materialise one main-shape witness; below 1.0 because "
                   "the last lines of each launch are still dirty in the 126 MB L2 when the kernel (and ncu's counter window) ends",
           "source": " + ".join(os.path.join("profiles", os.path.basename(f)) for f in argv) + " (ncu --set full, 1 witness per launch, main shape)",
           "kernels": ks}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "expand_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
