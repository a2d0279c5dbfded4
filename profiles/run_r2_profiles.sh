#!/bin/bash
# Round-2 final-state profiles (run on the GPU box from the repo root): launch list of one cfg4-shaped sweep, one
# `ncu --set full` capture of each hot kernel (raw + details pages), DRAM traffic per launch.
set -u
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_cfg4shape.csv python profiles/prof_workload.py 1 > $O/r2_prof.log 2>&1
cap() {  # name regex skip script args...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/r2_$name "$@" >> $O/r2_prof.log 2>&1
  ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > $O/r2_ncu_$name.csv 2>>$O/r2_prof.log
  ncu -i /tmp/r2_$name.ncu-rep --page details --csv 2>/dev/null | grep -i "Stall\|Throughput\|Pipe\|Occupancy\|Registers\|Duration\|Warp Cycles\|Issued\|DRAM\|L2" | head -80 > $O/r2_details_$name.csv
}
cap gemm "gemm_kernel<64" 40 python profiles/prof_workload.py 1
cap udt_steps_256 "udt_steps_kernel<32" 2 python profiles/prof_workload.py 1
cap udt_steps_192 "udt_steps_kernel<24" 2 python profiles/prof_workload.py 1
cap udt_formq4 udt_formq4_kernel 2 python profiles/prof_workload.py 1
cap update3 update3_kernel 3 python profiles/prof_workload.py 1
cap slice_steps slice_steps_kernel 2 python profiles/sweep_one.py cfg2
du -sh $O
