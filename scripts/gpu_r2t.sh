#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_sweep.sh ${1:-r2t} 6 | cut -c1-420
