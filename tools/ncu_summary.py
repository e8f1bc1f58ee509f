"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into one line per launch:
duration, DRAM bytes, DRAM / L2 throughput, occupancy, registers.  Used to fill profiles/."""
import csv
import io
import subprocess
import sys

COLS = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit",
    "lts__t_sector_hit_rate.pct": "l2_hit",
}


def to_num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return x * scale.get(unit, 1.0)


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    print("kernel | grid | dur_us | dram_rd_MB | dram_wr_MB | dram_GBs | dram_pct | l2_MB | l1_hit | l2_hit | occ_pct | tensor_pct | regs")
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("hgin::", "").replace("(anonymous namespace)::", "")
        vals = {}
        for m, short in COLS.items():
            if m in idx:
                vals[short] = to_num(r[idx[m]], units[idx[m]])
        dur = vals.get("dur") or 0.0
        rd, wr = vals.get("dram_rd") or 0.0, vals.get("dram_wr") or 0.0
        gbs = (rd + wr) / (dur * 1e-6) / 1e9 if dur else 0.0
        f = lambda k, s=1.0: ("%.1f" % (vals[k] / s)) if vals.get(k) is not None else "-"
        print(f"{name[:60]} | {r[idx['Grid Size']]} | {dur:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | "
              f"{f('dram_pct')} | {f('l2_bytes', 1e6)} | {f('l1_hit')} | {f('l2_hit')} | {f('occ_pct')} | {f('tensor_pct')} | {f('regs')}")


if __name__ == "__main__":
    main(sys.argv[1])
