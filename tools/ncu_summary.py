"""Cut the raw-page CSV of an `ncu --set full` capture of the window-GEMM launches of one step down to the columns the
design notes quote:  python tools/ncu_summary.py gpurun_out/r02_wconv_raw.csv > profiles/r02_ncu_wconv_full.csv"""
import csv
import sys

LAYERS = ["prior_network.1", "prior_network.2", "prior_network.3", "p_y_z_in.0 conv5 3->16", "p_y_z_in.1 k4s2 16->32",
          "p_y_z_in.2 k4s2 32->64", "p_y_z_in.3 k4s2 64->128"] + \
         ["res%d.%s" % (b, c) for b in range(4) for c in ("open", "close")] + \
         ["T12 128->64", "T13 64->32", "T14 32->16", "p_mu_out.0 conv7 16->8", "p_mu_out.1 conv5 8->1"]
COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    w = csv.writer(sys.stdout)
    w.writerow(["layer"] + [hdr[i] for i in idx])
    w.writerow(["-"] + [units[i] for i in idx])
    for name, r in zip(LAYERS, rows[2:]):
        w.writerow([name] + [r[i] for i in idx])


if __name__ == "__main__":
    main(sys.argv[1])
