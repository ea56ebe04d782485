"""Condense an `ncu --set full` capture of `python bench.py` into profiles/ncu_traffic.json: DRAM bytes (read + write) of one launch
of every step kernel, keyed by bench.py's kernel labels, stamped with the workload, frame count and the source hash of the library
that was profiled.  bench.py reports `roofline.traffic` from this table only when all three match the running build.

    python profiles/ncu_traffic.py gpurun_out/r02_prof_c3.ncu-rep --workload c3 --frames 4194304 --source-hash <hash> [--out profiles/ncu_traffic.json]
"""
import argparse
import csv
import json
import os
import subprocess

LABELS = (("prep_align_kernel", "fast_prep"), ("prep_feat_kernel", "fast_prep"), ("prep_transpose_kernel", "fast_prep"),
          ("pass1_kernel", "fast_pass1"), ("pass2_kernel", "fast_pass2"), ("jjt_kernel", "fast_jjt"), ("stats_kernel", "fast_stats"),
          ("align_tile_kernel", "align_fwd"), ("ae_fast_main", "ae_fast_main"), ("ae_fast_dw", "ae_fast_dw"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--frames", type=int, required=True)
    ap.add_argument("--source-hash", required=True)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"))
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    kernels, times = {}, {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        for pat, label in LABELS:
            if pat in name:
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                unit_r, unit_w = rows[1][hdr.index("dram__bytes_read.sum")], rows[1][hdr.index("dram__bytes_write.sum")]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                kernels[label] = rd * scale.get(unit_r, 1.0) + wr * scale.get(unit_w, 1.0)
                times[label] = float(d["gpu__time_duration.sum"])
    tab = {"workload": a.workload, "frames": a.frames, "source_hash": a.source_hash, "capture": os.path.basename(a.rep),
           "kernels": kernels, "gpu_time_under_ncu": times,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full --clock-control none); times under ncu are "
                   "cold-cache and serialised, not bench values"}
    with open(a.out, "w") as f:
        json.dump(tab, f, indent=1)
    print(json.dumps(tab, indent=1))


if __name__ == "__main__":
    main()
