mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest "$@" -q -s -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $? : $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt; }
run osp 500 tests/test_osp_gpu.py
run elem 300 tests/test_kernels_gpu.py -k "ln_modulate or patch_embed or qk_rmsnorm or flash"
grep -h "parity\]\|pipeline\]" gpurun_out/osp.log | grep -i "mma\|v120" | tail -n 40
grep -h "Error\|error\|FAILED" gpurun_out/osp.log gpurun_out/elem.log | head -20
