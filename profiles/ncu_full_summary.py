import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = [('Kernel Name','name'),('gpu__time_duration.sum','us'),('dram__bytes_read.sum','rdMB'),('dram__bytes_write.sum','wrMB'),
 ('launch__grid_size','grid'),('launch__block_size','blk'),('launch__registers_per_thread','regs'),
 ('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),('l1tex__t_sector_hit_rate.pct','l1hit'),('lts__t_sector_hit_rate.pct','l2hit'),
 ('smsp__inst_executed.sum','inst'),('sm__throughput.avg.pct_of_peak_sustained_elapsed','sm%'),
 ('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%'),('smsp__thread_inst_executed_per_inst_executed.ratio','thr/inst'),
 ('sm__inst_executed_pipe_fp64.sum','fp64')]
idx=[(h.index(w),n) for w,n in want if w in h]
print(' | '.join(n for _,n in idx)); print(' | '.join(rows[1][i] for i,_ in idx))
for r in rows[2:]:
    print(' | '.join(r[i][:40] for i,_ in idx))
