"""Key metrics of every launch in an .ncu-rep (`ncu --set full`), as text for profiles/."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: {len(rows) - 2} launch(es); ncu --set full --clock-control none (cold-cache, serialised replays)")
    for r in rows[2:]:
        print("---")
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w or (h.startswith(w) and h[len(w):] in ("", ".per_second", ".pct_of_peak_sustained_elapsed")):
                    print(f"{h}: {r[i]} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
