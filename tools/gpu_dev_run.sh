#!/bin/bash
# Dev helper for `gpurun`: runs each GPU test group in its own process (a trapped kernel poisons only its group),
# every group under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  echo "=== $name ===" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest "$@" -q -s -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $? : $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
for g in "$@"; do
  case $g in
    elem)  run elem 300 tests/test_kernels_gpu.py -k "ln_modulate or gate_residual or qk_rmsnorm" ;;
    short) run short 300 tests/test_kernels_gpu.py -k "attn_short" ;;
    gemm)  run gemm 300 tests/test_kernels_gpu.py -k "gemm" ;;
    flash) run flash 300 tests/test_kernels_gpu.py -k "attn_flash" ;;
    model) run model 600 tests/test_model_gpu.py ;;
    all)   run all 900 tests ;;
    *)     run "custom" 600 $g ;;
  esac
done
grep -h "parity\]\|PASS\|FAIL\|passed\|failed\|error\|watchdog" gpurun_out/*.log | tail -n 60
