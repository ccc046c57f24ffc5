#!/bin/bash
# First-contact diagnostics on a GPU box: each group in its own process (a trapped kernel poisons the
# CUDA context), each under its own timeout.  Output -> gpurun_out/diag_*.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_smi.log 2>&1
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/diag_$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/diag_summary.log; tail -5 gpurun_out/diag_$name.log; }
run elementwise python -m pytest tests/test_gpu_elementwise.py -q -x -m gpu
run conv_simt python -m pytest tests/test_gpu_conv.py -q -m gpu -k "simt or argument"
run conv_tc python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tc-f16"
run conv_tc_bf16 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tc-bf16"
run models_simt python -m pytest tests/test_gpu_models.py -q -m gpu -k "simt or verify or raises"
run models_tc python -m pytest tests/test_gpu_models.py -q -m gpu -k "fp16-tc or invariance"
