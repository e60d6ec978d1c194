#!/bin/bash
# Round-2 profile capture (run under gpurun, one GPU).  Every ncu command runs only after the same command line
# exited 0 without ncu.  Outputs: gpurun_out/r02p_*.{ncu-rep,csv,log}
set -u
B="python bench.py --no-cpu-baseline --no-e2e --no-workloads --warmup 3"
run() {  # tag, kernel regex, skip, bench args...
  tag=$1; rx=$2; skip=$3; shift 3
  $B "$@" > gpurun_out/r02p_${tag}_plain.json 2> gpurun_out/r02p_${tag}_plain.err &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r02p_$tag \
      $B "$@" > gpurun_out/r02p_${tag}_ncu.log 2>&1
  echo "$tag rc=$?"
}
# launch list of the headline command (cold-cache, serialised: shares only)
$B --workload c2 --steps 5 > gpurun_out/r02p_launches_plain.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02p_c2_launches.csv \
    $B --workload c2 --steps 5 > gpurun_out/r02p_launches_ncu.log 2>&1
$B --workload c5_grid --strides 20 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/r02p_c5grid_launches.csv \
    $B --workload c5_grid --strides 20 > gpurun_out/r02p_c5grid_launches_ncu.log 2>&1
run c2 "md_block.*f2" 8 --workload c2 --steps 3
run c2_nodump "md_block.*f2" 4 --workload c2_nodump --steps 3
run c3 "md_block.*f2" 4 --workload c3
run c4 md_block 4 --workload c4 --steps 3
run dump1 "md_block.*f2" 4 --workload dump1 --steps 3
run c5 mtd_block 20 --workload c5 --strides 25
run c5grid md_block_wide 30 --workload c5_grid --strides 20
run gridacc grid_accumulate 10 --workload c5_grid --strides 20
ls -la gpurun_out/r02p_* | head -40
