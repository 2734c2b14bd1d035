#!/usr/bin/env python
"""Reduce an `ncu --set full` capture of tools/prof_step.py into profiles/kernel_traffic.json: DRAM bytes actually moved
per sample by each library kernel (dram__bytes_read.sum + dram__bytes_write.sum of the captured launch / batch), which
bench.py reports as roofline.traffic (x its own batch).

    python tools/ncu_traffic.py gpurun_out/r2_prof_step2048.ncu-rep 2048 profiles/kernel_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

NAMES = [("pose_fwd_kernel", "pose_fwd"), ("blend_f16_panel_kernel", "blend_fwd"), ("lbs_fwd_warp_kernel", "lbs_fwd"),
         ("mask_kernel", "mask"), ("seg_fwd_kernel", "seg_fwd"), ("seg_bwd_kernel", "seg_bwd"),
         ("lbs_bwd_sampled_kernel", "lbs_bwd_vertex"), ("split3_gemm_kernel<112", "blend_bwd"), ("pose_bwd_kernel", "pose_bwd")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, batch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    per = {}
    for r in rows[2:]:
        for pat, name in NAMES:
            if pat in r[ik]:
                b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
                per.setdefault(name, []).append(b / batch)
    doc = {"source": "%s: ncu --set full --clock-control none, tools/prof_step.py --batch %d (dram__bytes_read.sum + "
                     "dram__bytes_write.sum per launch / batch; mean over the captured launches of each kernel)" % (rep, batch),
           "batch": batch, "bytes_per_sample": {k: sum(v) / len(v) for k, v in per.items()}}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
