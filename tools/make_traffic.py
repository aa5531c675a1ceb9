"""profiles/traffic.json from `ncu --set full` captures: measured DRAM bytes (read + write) per launch of the kernels the
bench line may name as dominant.   usage: python tools/make_traffic.py ROUND msda.ncu-rep [gemm.ncu-rep]
(bench.py reads the file for `roofline.traffic`; regenerate it every round, the capture is part of tools/gpu_profile_round.sh)"""
import csv
import io
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAGS = {"msda_bwd_kernel": "poet_msda_bwd[Lq=1600]", "msda_bwd_tile_kernel": "poet_msda_bwd[Lq=1600]", "msda_bwd_shared_kernel": "poet_msda_bwd[Lq=1600]",
        "msda_fwd_slab_kernel": "poet_msda_fwd[Lq=1600]"}


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        yield r[ik], float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]


def main():
    rnd, reps = sys.argv[1], sys.argv[2:]
    per = {}
    for rep in reps:
        for name, b in launches(rep):
            for key, tag in TAGS.items():
                if key in name:
                    per.setdefault(tag, []).append(b)
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (median over the captured launches) from "
                       "`ncu --set full --clock-control none` of one eager cfg2 step; keyed by bench.py kernel tag",
           "_source": f"{rnd}: " + ", ".join(os.path.basename(r) for r in reps) + time.strftime(" (%Y-%m-%d)")}
    for tag, v in per.items():
        out[tag] = int(statistics.median(v))
    print(json.dumps(out, indent=1))        # redirect into profiles/traffic.json (the GPU box's tree does not travel back)


if __name__ == "__main__":
    main()
