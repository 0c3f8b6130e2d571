"""Summarise an `ncu --set full` report (run where ncu is installed, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/r02_tc_wiki.ncu-rep profiles/r02_ncu_tc_wiki.txt [profiles/r02_traffic.json key1 key2 ...]

Writes one line per captured launch (grid, duration, DRAM bytes, tensor-pipe / warp activity, registers) and, if a JSON
path and keys are given, the per-launch DRAM traffic of the first launches under those keys (read by bench.py for
`roofline.traffic`)."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block")]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


lines, traffic = [], []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    parts = []
    for c, name in COLS:
        if c in idx:
            v = r[idx[c]]
            if c == "Kernel Name":
                v = v.split("(")[0]
            parts.append(f"{name}={v} {units[idx[c]]}".strip())
    lines.append("  ".join(parts))
    traffic.append(to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
                   + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]))
open(out, "w").write(f"# ncu --set full summary of {rep} ({len(lines)} launches; cold cache, serialised)\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
if len(sys.argv) > 4:
    jp, keys = sys.argv[3], sys.argv[4:]
    try:
        d = json.load(open(jp))
    except Exception:
        d = {}
    for k, t in zip(keys, traffic):
        d[k] = {"traffic_bytes": int(t), "source": f"ncu --set full, {rep.split('/')[-1]}, per launch (cold)"}
    json.dump(d, open(jp, "w"), indent=1)
