"""Per-kernel DRAM bytes, duration and DMMA-pipe activity from an `ncu --page raw --csv` export of tools/profile.sh
    python tools/kernel_traffic.py gpurun_out/r01d_prof_raw.csv "<source note>" > profiles/r01_kernel_traffic.json
bench.py reads the result for `roofline.traffic`."""
import csv
import json
import sys

KERNELS = ("cond_fwd_a", "cond_fwd_b", "mc_pass", "syrk", "cond_bwd_a", "cond_bwd_b")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "%": 1}
DMMA = "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"


def main(path, note):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(row, m):
        return float(row[ix[m]].replace(",", "")) * SCALE.get(units[ix[m]], 1)

    agg = {}
    for row in data:
        k = next((k for k in KERNELS if k in row[ix["Kernel Name"]]), None)
        if k is None:
            continue
        d = agg.setdefault(k, {"bytes": [], "ms": [], "dmma": []})
        d["bytes"].append(val(row, "dram__bytes_read.sum") + val(row, "dram__bytes_write.sum"))
        d["ms"].append(val(row, "gpu__time_duration.sum"))
        d["dmma"].append(val(row, DMMA) if DMMA in ix else 0.0)
    mean = lambda v: sum(v) / len(v)
    json.dump({"source": note, "points_per_launch": 1048576,
               "per_launch_dram_bytes": {k: mean(v["bytes"]) for k, v in agg.items()},
               "per_launch_ms_under_ncu": {k: mean(v["ms"]) for k, v in agg.items()},
               "dmma_pipe_pct_of_peak_sustained_active": {k: mean(v["dmma"]) for k, v in agg.items()}}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
