mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest "$@" -q -s -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $? : $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt; }
run pe 200 tests/test_kernels_gpu.py -k "patch_embed or qk_rmsnorm"
run vch 400 tests/test_vchitect_gpu.py
run osp 400 tests/test_osp_gpu.py
run latte 400 tests/test_latte_gpu.py
timeout 400 python bench.py --workload vchitect_2b_40f_288x480_100step --steps 3 --warmup 3 > gpurun_out/bench_vchitect.json 2> gpurun_out/bench_vchitect.err
echo "bench_vchitect exit $? : $(tail -c 700 gpurun_out/bench_vchitect.json)" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench_vchitect.err
grep -h "parity\]\|pipeline\]" gpurun_out/vch.log gpurun_out/osp.log | tail -n 50
