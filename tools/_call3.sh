mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name --format=csv | tee -a gpurun_out/summary.txt
run() { local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest "$@" -q -s -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $? : $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt; }
run dsp 420 tests/test_dsp_gpu.py
run spmodels 420 tests/test_sp_models_gpu.py
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench_n2 exit $? : $(tail -c 1200 gpurun_out/bench_n2.json | cut -c1-500)" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench_n2.err
grep -h "parity\]" gpurun_out/dsp.log gpurun_out/spmodels.log | tail -n 20
