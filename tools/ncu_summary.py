"""Print the handful of ncu metrics DESIGN.md / profiles/ quote from a .ncu-rep (run where ncu is installed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
    'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print('## %s' % rep)
        for r in rows[2:]:
            print('# kernel: %s' % r[ix['Kernel Name']][:100])
            for k in KEYS:
                if k in ix:
                    print('%-75s %14s %s' % (k, r[ix[k]], units[ix[k]]))


if __name__ == '__main__':
    main()
