#!/bin/bash
# second pass of profiles/capture_r02.sh: the packed kernels (ncu's -k matches the base name md_block)
set -u
B="python bench.py --no-cpu-baseline --no-e2e --no-workloads --warmup 3"
run() {  # tag, skip, count, bench args...
  tag=$1; skip=$2; cnt=$3; shift 3
  $B "$@" > gpurun_out/r02p_${tag}_plain.json 2> gpurun_out/r02p_${tag}_plain.err &&
  ncu --set full --clock-control none --import-source on -k regex:md_block -s $skip -c $cnt -o gpurun_out/r02p_$tag \
      $B "$@" > gpurun_out/r02p_${tag}_ncu.log 2>&1
  echo "$tag rc=$?"
}
run c2 8 2 --workload c2 --steps 3
run c2_nodump 4 1 --workload c2_nodump --steps 3
run c3 4 1 --workload c3
run dump1 4 1 --workload dump1 --steps 3
ls -la gpurun_out/r02p_*.ncu-rep
