#!/usr/bin/env python
"""Per-family DRAM traffic from an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch).
usage: python tools/launch_traffic.py gpurun_out/launches.csv profiles/r01_traffic.json"""
import csv
import json
import re
import sys

FAMILIES = (("dense_layer", "dense_layer"), ("conv1x1_persist", "conv1x1"), ("conv3x3_roll", "conv3x3"), ("conv3x3_rows", "conv3x3"),
            ("conv_gemm", "pool1x1"), ("stem", "stem"), ("linear", "linear"), ("gemm_tma", "fc_gemm"), ("split_bf16", "fc_gemm"), ("sg_render", "render"), ("head_pool", "head_pool"))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    launches = {}
    for r in csv.reader(open(src)):
        if len(r) > 10 and r[0].isdigit():
            launches.setdefault(int(r[0]), {"name": r[4]})[r[-3]] = float(r[-1])
    # keep exactly one step: from the first stem launch up to (not including) the next one
    order = sorted(launches)
    stems = [i for i in order if "stem" in launches[i]["name"]]
    if len(stems) >= 2:
        order = [i for i in order if stems[0] <= i < stems[1]]
    fam = {}
    for i in order:
        l = launches[i]
        name = next((f for key, f in FAMILIES if key in l["name"]), "other")
        if name == "dense_layer" and re.search(r"dense_layer_kernel<\d, 1[,>]", l["name"].replace("(bool)", "").replace("(int)", "")):
            name = "pool1x1"                                  # dense_layer_kernel<SPLIT, POOL = true> is transition1 (eml_conv_forward POOL2)
        f = fam.setdefault(name, {"launches": 0, "time_ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        f["launches"] += 1
        f["time_ms"] += l.get("gpu__time_duration.sum", 0.0) / 1e6
        f["dram_read_bytes"] += l.get("dram__bytes_read.sum", 0.0)
        f["dram_write_bytes"] += l.get("dram__bytes_write.sum", 0.0)
    total = sum(f["time_ms"] for f in fam.values())
    for f in fam.values():
        f["share_of_step"] = round(f["time_ms"] / total, 4)
        f["traffic_bytes_per_launch"] = round((f["dram_read_bytes"] + f["dram_write_bytes"]) / f["launches"])
        f["time_ms"] = round(f["time_ms"], 3)
    out = {"source": src, "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one bench step (B=256)",
           "step_ms_under_ncu": round(total, 3), "families": fam}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
