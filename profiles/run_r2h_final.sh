#!/bin/bash
# Round-2 final-state profiles (run on the GPU box from the repo root): launch list of one cfg4-shaped sweep, one
# `ncu --set full` capture of each hot kernel (raw + details pages; kernels selected by name and launch index).
set -u
O=gpurun_out
P=r2h
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${P}_launches_cfg4shape.csv python profiles/prof_workload.py 1 > $O/${P}_prof.log 2>&1
cap() {  # name kernel skip script args...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k $k -s $skip -c 1 -f -o /tmp/${P}_$name "$@" >> $O/${P}_prof.log 2>&1
  ncu -i /tmp/${P}_$name.ncu-rep --page raw --csv > $O/${P}_ncu_$name.csv 2>>$O/${P}_prof.log
  ncu -i /tmp/${P}_$name.ncu-rep --page details --csv 2>/dev/null | grep -i "Stall\|Throughput\|Pipe\|Occupancy\|Registers\|Duration\|Warp Cycles\|Issued\|DRAM\|L2" | head -80 > $O/${P}_details_$name.csv
}
# launch order of one UDT call: steps 256, 192, 128, 96, 64, wy_t, formq4
cap gemm gemm_kernel 60 python profiles/prof_workload.py 1
cap udt_steps_256 udt_steps_kernel 10 python profiles/prof_workload.py 1
cap udt_steps_192 udt_steps_kernel 11 python profiles/prof_workload.py 1
cap udt_formq4 udt_formq4_kernel 2 python profiles/prof_workload.py 1
cap update3 update3_kernel 3 python profiles/prof_workload.py 1
cap slice_steps slice_steps_kernel 2 python profiles/sweep_one.py cfg2
for n in gemm udt_steps_256 udt_steps_192 udt_formq4 update3 slice_steps; do head -c 600 $O/${P}_ncu_$n.csv | tail -c 200; echo; done
du -sh $O
