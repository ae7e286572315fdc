#!/bin/bash
mkdir -p gpurun_out
for CFG in water_23k protein_92k; do
  timeout 420 python baseline/ref_numba_cuda.py --config $CFG --evals 3 > gpurun_out/ref_numba_cuda_$CFG.json 2> gpurun_out/ref_numba_cuda_$CFG.err
  echo "$CFG rc=$?"; tail -c 700 gpurun_out/ref_numba_cuda_$CFG.err; tail -1 gpurun_out/ref_numba_cuda_$CFG.json
done
