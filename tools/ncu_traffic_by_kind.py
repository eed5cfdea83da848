#!/usr/bin/env python
"""Reduce an ncu launch list (CSV, one row per launch and metric) of `bench.py` to DRAM bytes and time per kernel
FAMILY of the forward plan.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 600 \
        --csv --log-file launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-latency \
        --dump-profile prof.json
    python tools/ncu_traffic_by_kind.py launches.csv prof.json out.json

The plan's ops (prof.json: name, kind, ..., in launch order) are matched to the launches of each captured forward in
order — a pointwise GEMM and a 3x3 GEMM are the same SASS kernel but different families — so that
`roofline.traffic` in bench.py is the traffic of the dominant FAMILY only.  The first forward after process start
(cold instruction / descriptor caches) is skipped when more than one was captured.
"""
from __future__ import annotations

import csv
import json
import sys

KERNEL_OF_KIND = {
    "stem_pool": "stem_pool_kernel", "dwconv3x3": "dwconv3x3", "pw_tcgen05": "tc_gemm_kernel",
    "conv3x3_tcgen05": "tc_gemm_kernel", "pw_decode_tcgen05": "tc_gemm_kernel", "dwpw_tcgen05": "dwpw_tc_kernel",
    "resample_add": "resample_add_kernel", "nms": "nms_", "decode": "decode_level_kernel",
    "pw_ffma": "gemm_ffma_kernel", "conv3x3_ffma": "gemm_ffma_kernel", "interleave_copy": "interleave_copy_kernel",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def read_launches(path):
    hdr, launches = None, {}
    for r in csv.reader(open(path, newline="")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        lid = int(d["ID"])
        e = launches.setdefault(lid, {"name": d["Kernel Name"], "dram": 0.0, "time_s": 0.0, "pct": {}})
        try:
            v = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
        except ValueError:
            continue
        if d["Metric Name"].startswith("dram__bytes"):
            e["dram"] += v
        elif d["Metric Name"].startswith("gpu__time_duration"):
            e["time_s"] += v
        elif d["Metric Unit"] == "%":
            e["pct"][d["Metric Name"]] = v
    return [launches[k] for k in sorted(launches)]


def main():
    launches = read_launches(sys.argv[1])
    ops = json.load(open(sys.argv[2]))["per_launch_last_rep"]
    starts = [i for i, l in enumerate(launches) if "stem_pool_kernel" in l["name"]]
    forwards = []
    for s in starts:
        i, rows, ok = s, [], True
        for name, kind, *_ in ops:
            want = KERNEL_OF_KIND.get(kind, kind)
            if kind == "nms":                     # one op = several kernels
                got = []
                while i < len(launches) and want in launches[i]["name"]:
                    got.append(launches[i]); i += 1
                if not got:
                    ok = False
                    break
                tt = sum(g["time_s"] for g in got) or 1.0
                pct = {k: sum(g["pct"].get(k, 0.0) * g["time_s"] for g in got) / tt for k in got[0]["pct"]}
                rows.append((name, kind, sum(g["dram"] for g in got), sum(g["time_s"] for g in got), len(got), pct))
                continue
            if i >= len(launches) or want not in launches[i]["name"]:
                ok = False
                break
            rows.append((name, kind, launches[i]["dram"], launches[i]["time_s"], 1, launches[i]["pct"]))
            i += 1
        if ok:
            forwards.append(rows)
    if not forwards:
        raise SystemExit("no complete forward found in the launch list")
    use = forwards[1:] if len(forwards) > 1 else forwards
    per_kind, total_t = {}, 0.0
    for rows in use:
        for name, kind, dram, t, n, pct in rows:
            d = per_kind.setdefault(kind, {"dram": 0.0, "time_s": 0.0, "launches": 0, "kernels": 0, "pct": {}})
            d["dram"] += dram; d["time_s"] += t; d["launches"] += 1; d["kernels"] += n
            for k, v in pct.items():
                d["pct"][k] = d["pct"].get(k, 0.0) + v * t          # time-weighted
            total_t += t
    out = {"source": sys.argv[1], "forwards_captured": len(forwards), "forwards_used": len(use),
           "note": "per-launch ncu times are cold-cache and serialised: compare SHARES with the CUDA-event numbers",
           "per_kind": {k: {"dram_bytes_per_launch": v["dram"] / v["launches"], "dram_bytes_per_step": v["dram"] / len(use),
                            "launches_per_step": v["launches"] / len(use), "time_share": v["time_s"] / total_t,
                            "pct_time_weighted": {k: round(x / v["time_s"], 2) for k, x in sorted(v["pct"].items())}}
                        for k, v in sorted(per_kind.items())},
           "dram_bytes_per_step": sum(v["dram"] for v in per_kind.values()) / len(use)}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out["per_kind"], indent=1))
    print("DRAM bytes per step: %.3f GB" % (out["dram_bytes_per_step"] / 1e9))


if __name__ == "__main__":
    main()
