"""print selected raw metrics of an .ncu-rep (offline)"""
import csv, sys, subprocess, io
rep = sys.argv[1]
pat = sys.argv[2:] or ['gpu__time_duration.sum','launch__registers','launch__shared_mem_per_block_dynamic','launch__occupancy_limit','launch__waves','sm__warps_active.avg.pct','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct','lts__t_bytes.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_active.avg','gpc__cycles_elapsed.max','sm__inst_executed.sum','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','smsp__average_warps_issue_stalled','smsp__warp_issue_stalled','sm__pipe_tensor','smsp__issue_active.avg.pct','sm__inst_executed_pipe','lts__throughput','l1tex__throughput','smsp__inst_executed_pipe_xu','sm__cycles_elapsed']
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[-1]
print(r[hdr.index('Kernel Name')][:70])
for i, h in enumerate(hdr):
    if any(p in h for p in pat):
        try:
            v = float(r[i].replace(',', ''))
            if v == 0: continue
        except ValueError:
            pass
        print('  %-90s %s %s' % (h, r[i], units[i]))
