"""Regenerate profiles/ncu_traffic.json (the per-launch DRAM bytes bench.py copies into roofline.traffic) from a
tools/ncu_summary.py summary of ONE Runge-Kutta stage of the 512^3 TGV block (`ncu --set full` of
`bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --parity-n 0`, launch window = the last stage).
usage: python tools/ncu_traffic.py profiles/r02_stage_v3_ncu.txt > profiles/ncu_traffic.json

The launch order of a stage on the fused path (astr_gpu_rk_stage, api.cu) names the kernels: filter i, j, k ->
gradient i, j, k -> k_visc_flux -> divergence i, j, k -> k_rk_update."""
import json
import re
import sys

ORDER = ["filter_i", "filter_j", "filter_k", "grad_i", "grad_j", "grad_k", "visc", "div_i", "div_j", "div_k", "rk"]
FIELDS = {"filter": 5, "grad": 4, "div": 5}
PTS = 513 ** 3
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(path):
    launches = []
    for line in open(path):
        if line.startswith("== "):
            launches.append({"name": line[3:].strip()})
        m = re.match(r"\s+(dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__time_duration\.sum)\s+([\d.,]+)\s+(\S+)", line)
        if m and launches:
            launches[-1][m.group(1)] = float(m.group(2).replace(",", "")) * SCALE.get(m.group(3), 1.0)
    # the sweeps and the two large pointwise kernels only (halo wraps etc. are not in the capture window)
    big = [l for l in launches if re.search(r"sweep2i?_kernel|k_visc_flux|k_rk_update", l["name"])]
    assert len(big) == len(ORDER), f"expected one stage = {len(ORDER)} launches, found {len(big)}"
    out = {"source": f"ncu --set full --clock-control none, one stage of the 512^3 TGV block; see {path}",
           "n": 512, "bytes_per_launch": {}, "read_bytes_per_launch": {}, "algorithmic_bytes_per_launch": {},
           "kernel": {}, "ms_under_ncu": {}}
    for key, l in zip(ORDER, big):
        fam = key.split("_")[0]
        assert ("sweep2" in l["name"]) == (fam in FIELDS), (key, l["name"])
        out["bytes_per_launch"][key] = l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
        out["read_bytes_per_launch"][key] = l["dram__bytes_read.sum"]
        out["algorithmic_bytes_per_launch"][key] = (FIELDS[fam] * 16 if fam in FIELDS else {"visc": 47 * 8, "rk": 37 * 8}[key]) * PTS
        out["kernel"][key] = l["name"]
        out["ms_under_ncu"][key] = l["gpu__time_duration.sum"]
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
