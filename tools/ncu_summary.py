"""Summarise .ncu-rep captures (read with `ncu -i ... --page raw --csv`, no GPU needed) into the
small text files committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/r1_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed.sum",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"== {rep}: unreadable")
            continue
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            print(f"== {rep}  kernel: {d.get('Kernel Name', '?')[:100]}")
            for k in KEYS:
                if k in d and d[k] not in ("", "n/a"):
                    print(f"   {k:100s} {d[k]:>16s} {u[k]}")
            try:
                rd = float(d.get("dram__bytes_read.sum", "0").replace(",", "") or 0)
                wr = float(d.get("dram__bytes_write.sum", "0").replace(",", "") or 0)
                print(f"   traffic = dram read + write = {rd + wr:.3f} (unit of the two rows above)")
            except ValueError:
                pass


if __name__ == "__main__":
    main()
