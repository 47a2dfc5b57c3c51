"""Summarises an .ncu-rep (or a launch-list csv) into a small markdown table under profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "sm__cycles_elapsed.max", "lts__t_bytes.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in KEYS if k in hdr]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n| kernel | " + " | ".join(
            f"{hdr[c]} [{units[c]}]" for c in cols) + " |\n|" + "---|" * (len(cols) + 1) + "\n")
        for r in rows[2:]:
            f.write(f"| {r[ki][:40]} | " + " | ".join(r[c] for c in cols) + " |\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
