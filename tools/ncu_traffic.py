#!/usr/bin/env python
"""Per-kernel summary and DRAM traffic from an `ncu --set full` report (run here, no GPU needed):

    python tools/ncu_traffic.py gpurun_out/r02_x_full.ncu-rep --summary profiles/r02_ncu_summary_x.txt \
        [--traffic profiles/r02_traffic.json --B 4096 --P 50 --n 1024]

--summary: one line per profiled launch (duration, DRAM bytes read / written, DRAM and issue utilisation, registers,
warp instructions, shared-memory wavefronts, L1 data-pipe utilisation, top stall reasons).
--traffic: `dram__bytes_read.sum + dram__bytes_write.sum` per launch, averaged per kernel family, in the form bench.py reads
for `roofline.traffic` (keys contract / step_fwd / step_bwd); the launches of the first (k = 0) and last level are kept in
the average -- they move fewer bytes, as they do in the step."""
import argparse
import csv
import json
import subprocess
import sys
from collections import OrderedDict

M = {
    "t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread", "inst": "smsp__inst_executed.sum", "smem_wf": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "grid": "launch__grid_size", "block": "launch__block_size",
}


def family(name):
    if "contract_f16" in name or "contract_tc" in name:
        return "contract"
    if "level_fwd" in name:
        return "step_fwd"
    if "level_bwd" in name:
        return "step_bwd"
    return None


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--summary")
    ap.add_argument("--traffic")
    ap.add_argument("--B", type=int); ap.add_argument("--P", type=int); ap.add_argument("--n", type=int)
    ap.add_argument("--note", default="")
    o = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", o.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(v) for k, v in M.items() if v in hdr}
    stalls = [(i, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    lines, fam = [], OrderedDict()
    for r in data:
        name = r[hdr.index("Kernel Name")]
        g = lambda k: r[col[k]] if k in col else "-"
        rd, wr = to_bytes(g("rd"), units[col["rd"]]), to_bytes(g("wr"), units[col["wr"]])
        top = sorted(((float(r[i]), s) for i, s in stalls), reverse=True)[:3]
        lines.append(f"{name[:58]:58s} {g('t'):>9s} {units[col['t']]:<3s} dram rd {rd / 1e9:6.3f} GB wr {wr / 1e9:6.3f} GB  dram {float(g('dram')):5.1f}%  issue {float(g('issue')):5.1f}%  "
                     f"tensor {float(g('tensor')):5.1f}%  L1pipe {float(g('l1')):5.1f}%  warps {float(g('warps')):5.1f}%  regs {g('regs'):>3s}  grid {g('grid')}x{g('block')}  "
                     f"inst {float(g('inst')) / 1e6:7.1f} M  smem wf {float(g('smem_wf')) / 1e6:6.1f} M  stalls " + ", ".join(f"{s} {v:.2f}" for v, s in top))
        f = family(name)
        if f:
            fam.setdefault(f, []).append(rd + wr)
    if o.summary:
        with open(o.summary, "w") as fh:
            fh.write(f"# {o.rep}: ncu --set full --clock-control none, one line per profiled launch (cold-cache, serialised).  {o.note}\n")
            fh.write("\n".join(lines) + "\n")
    else:
        print("\n".join(lines))
    if o.traffic:
        out = {"source": f"tools/ncu_traffic.py on {o.rep} ({o.note})".strip(), "B": o.B, "P": o.P, "n": o.n,
               "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in fam.items()}, "launches_averaged": {k: len(v) for k, v in fam.items()}}
        json.dump(out, open(o.traffic, "w"), indent=1)
        print(json.dumps(out["dram_bytes_per_launch"]))


if __name__ == "__main__":
    main()
