"""Summarises .ncu-rep captures (ncu --set full) into the few metrics DESIGN.md / bench.py quote.
  python tools/ncu_summary.py title1=path1.ncu-rep [title2=path2.ncu-rep ...] > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    for arg in sys.argv[1:]:
        title, path = arg.rsplit("=", 1)
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("## %s: could not read %s" % (title, path))
            continue
        hdr, units = rows[0], rows[1]
        print("## %s  (ncu --set full --clock-control none --import-source on; %s)" % (title, path.split("/")[-1]))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("  Kernel Name".ljust(70), d.get("Kernel Name", "?")[:90])
            for k in WANT:
                if k in d:
                    print(("  " + k).ljust(70), d[k], dict(zip(hdr, units)).get(k, ""))
            print()


if __name__ == "__main__":
    main()
