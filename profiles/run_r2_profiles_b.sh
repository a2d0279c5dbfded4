#!/bin/bash
# second pass: the captures whose kernel regex did not match the first time (demangled names carry "(int)")
set -u
O=gpurun_out
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/r2_$name "$@" >> $O/r2_prof.log 2>&1
  ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > $O/r2_ncu_$name.csv 2>>$O/r2_prof.log
  ncu -i /tmp/r2_$name.ncu-rep --page details --csv 2>/dev/null | grep -i "Stall\|Throughput\|Pipe\|Occupancy\|Registers\|Duration\|Warp Cycles\|Issued\|DRAM\|L2" | head -80 > $O/r2_details_$name.csv
}
cap gemm gemm_kernel 60 python profiles/prof_workload.py 1
cap udt_steps_256 udt_steps_kernel 8 python profiles/prof_workload.py 1
cap udt_steps_192 udt_steps_kernel 9 python profiles/prof_workload.py 1
head -c 300 $O/r2_ncu_gemm.csv | tail -c 120
