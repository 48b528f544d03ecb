#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one block per launch."""
import csv, subprocess, sys
WANT = [
 ('time_us','gpu__time_duration.sum'),('dram_rd_GB','dram__bytes_read.sum'),('dram_wr_GB','dram__bytes_write.sum'),
 ('dram_pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),('sm_pct','sm__throughput.avg.pct_of_peak_sustained_elapsed'),
 ('regs','launch__registers_per_thread'),('warps_active_pct','sm__warps_active.avg.pct_of_peak_sustained_active'),
 ('warp_inst','smsp__inst_executed.sum'),('issue_active_pct','smsp__issue_active.avg.pct_of_peak_sustained_active'),
 ('fma_pipe_pct','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
 ('fmaheavy_pct','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active'),
 ('lsu_pct','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
 ('smem_conflicts','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
 ('l1_hit_pct','l1tex__t_sector_hit_rate.pct'),('l2_hit_pct','lts__t_sector_hit_rate.pct'),
 ('l2_rd_sectors','lts__t_sectors_op_read.sum'),('l2_wr_sectors','lts__t_sectors_op_write.sum'),
 ('grid','launch__grid_size'),('occ_lim_smem','launch__occupancy_limit_shared_mem'),('occ_lim_regs','launch__occupancy_limit_registers'),
 ('stall_long_sb','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
 ('stall_short_sb','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
 ('stall_barrier','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
 ('stall_mio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio'),
 ('stall_lg','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio'),
 ('stall_wait','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
 ('stall_math','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio'),
 ('stall_notsel','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio'),
 ('stall_dispatch','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio'),
 ('stall_nc','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'),
]
def main(path, pixels=None):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h:i for i,h in enumerate(hdr)}
    for r in rows[2:]:
        print('###', r[idx['Kernel Name']][:150])
        line=[]
        for k,m in WANT:
            if m in idx:
                v=r[idx[m]]
                try: v='%.4g'%float(v)
                except: pass
                line.append('%s=%s%s'%(k,v,'' if units[idx[m]] in ('','inst','cycle','block','register/thread','%','sector') else units[idx[m]]))
        print('   '+'  '.join(line))
if __name__=='__main__': main(*sys.argv[1:])
