#!/bin/bash
# Does the L2 fetch granularity change what the gathers of the write pass pull from DRAM?  (10 % selectivity, write_kernel)
set -x
cd /root/repo
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for g in default 32 128; do
  if [ $g = default ]; then unset MBC_L2_FETCH; else export MBC_L2_FETCH=$g; fi
  ENGINES=twopass timeout -s KILL 300 ncu --metrics $M --clock-control none -k regex:write_kernel -s 3 -c 1 --csv --log-file gpurun_out/l2_$g.csv python scripts/bench_engines.py 100000000 1 0.1 > gpurun_out/l2_$g.log 2>&1
  grep -h "L2 fetch" gpurun_out/l2_$g.log | head -1
  tail -6 gpurun_out/l2_$g.csv | cut -d, -f5,13-
done
