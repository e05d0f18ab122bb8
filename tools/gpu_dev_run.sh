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
    short) run short 300 tests/test_kernels_gpu.py -k "attn_short or patch_embed or temporal" ;;
    gemm)  run gemm 300 tests/test_kernels_gpu.py -k "gemm and not cta_pair" ;;
    gemm2) run gemm2 300 tests/test_kernels_gpu.py -k "cta_pair or single_cta or fused_residual" ;;
    pipe)  run pipe 300 tests/test_pipeline_gpu.py ;;
    bench2sm) timeout 900 python bench.py --opt gemm_2sm=1 $BENCH_ARGS > gpurun_out/bench2sm.json 2> gpurun_out/bench2sm.err
           echo "bench2sm exit $? : $(tail -c 300 gpurun_out/bench2sm.json)" | tee -a gpurun_out/summary.txt ;;
    flash) run flash 300 tests/test_kernels_gpu.py -k "attn_flash" ;;
    model) run model 600 tests/test_model_gpu.py ;;
    all)   run all 900 tests ;;
    dsp)   run dsp 600 tests/test_dsp_gpu.py ;;
    spmodels) run spmodels 600 tests/test_sp_models_gpu.py ;;
    mcogx*) n=${g#mcogx}; echo "=== bench cogvideox N=$n ($COGX_ARGS) ===" | tee -a gpurun_out/summary.txt
           timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29535 \
              bench.py --gpus $n --workload cogvideox_2b_49f_480x720_50step $COGX_ARGS > gpurun_out/benchcogx_n$n.json 2> gpurun_out/benchcogx_n$n.err
           echo "exit $? : $(tail -c 1500 gpurun_out/benchcogx_n$n.json | cut -c1-400)" | tee -a gpurun_out/summary.txt
           tail -n 8 gpurun_out/benchcogx_n$n.err ;;
    refgpu) timeout 600 python tests/bench_reference_gpu.py > gpurun_out/refgpu.log 2>&1; echo "refgpu exit $? $(tail -n 1 gpurun_out/refgpu.log | cut -c1-300)" | tee -a gpurun_out/summary.txt ;;
    mmab) timeout 120 tools/_bin/mma_microbench > gpurun_out/mma_microbench.txt 2>&1; echo "mmab exit $?" | tee -a gpurun_out/summary.txt ;;
    atrace) timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.log 2>&1; echo "atrace exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/attn_trace.log ;;
    kbattn) KB_ONLY=attn timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench_attn.log 2>&1; echo "kbattn exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/kernel_bench_attn.log ;;
    kbelem) KB_ONLY=elem timeout 300 python tools/kernel_bench.py > gpurun_out/kernel_bench_elem.log 2>&1; echo "kbelem exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/kernel_bench_elem.log ;;
    kernels) run kernels 900 tests/test_kernels_gpu.py ;;
    cogx) run cogx 600 tests/test_cogvideox_gpu.py ;;
    latte) run latte 600 tests/test_latte_gpu.py ;;
    vch) run vch 600 tests/test_vchitect_gpu.py ;;
    osp) run osp 600 tests/test_osp_gpu.py ;;
    benchvch) timeout 600 python bench.py --workload vchitect_2b_40f_288x480_100step $BENCH_ARGS > gpurun_out/bench_vchitect.json 2> gpurun_out/bench_vchitect.err
           echo "benchvch exit $? : $(tail -c 400 gpurun_out/bench_vchitect.json)" | tee -a gpurun_out/summary.txt ;;
    mmabench) timeout 300 python tools/attn_mma_bench.py > gpurun_out/attn_mma_bench.json 2> gpurun_out/attn_mma_bench.err
           echo "mmabench exit $? : $(tail -c 500 gpurun_out/attn_mma_bench.json)" | tee -a gpurun_out/summary.txt ;;
    benchcogx) timeout 900 python bench.py --workload cogvideox_2b_49f_480x720_50step $BENCH_ARGS > gpurun_out/benchcogx.json 2> gpurun_out/benchcogx.err
           echo "benchcogx exit $? : $(tail -c 500 gpurun_out/benchcogx.json)" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/benchcogx.err ;;
    kbench) timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/kernel_bench.log | tail -12 ;;
    mbenchnccl*) n=${g#mbenchnccl}; echo "=== bench N=$n (NCCL a2a) ===" | tee -a gpurun_out/summary.txt
           VSB_DSP_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 \
              bench.py --gpus $n $BENCH_ARGS > gpurun_out/bench_nccl_n$n.json 2> gpurun_out/bench_nccl_n$n.err
           echo "exit $? : $(tail -c 1500 gpurun_out/bench_nccl_n$n.json | cut -c1-300)" | tee -a gpurun_out/summary.txt ;;
    mbench*) n=${g#mbench}; echo "=== bench N=$n ===" | tee -a gpurun_out/summary.txt
           timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
              bench.py --gpus $n $BENCH_ARGS > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
           echo "exit $? : $(tail -c 1500 gpurun_out/bench_n$n.json | cut -c1-400)" | tee -a gpurun_out/summary.txt
           tail -n 8 gpurun_out/bench_n$n.err ;;
    smoke) echo "=== smoke ===" | tee -a gpurun_out/summary.txt
           timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
           echo "exit $? : $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt ;;
    bench) echo "=== bench ===" | tee -a gpurun_out/summary.txt
           timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
           echo "exit $? : $(tail -c 600 gpurun_out/bench.json)" | tee -a gpurun_out/summary.txt
           tail -n 5 gpurun_out/bench.err ;;
    bench240) timeout 600 python bench.py --workload opensora_240p_51f_30step --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench240.json 2> gpurun_out/bench240.err
           echo "bench240 exit $? : $(tail -c 400 gpurun_out/bench240.json)" | tee -a gpurun_out/summary.txt ;;
    benchpab) timeout 600 python bench.py --pab --no-cpu-baseline $BENCH_ARGS > gpurun_out/benchpab.json 2> gpurun_out/benchpab.err
           echo "benchpab exit $? : $(tail -c 400 gpurun_out/benchpab.json)" | tee -a gpurun_out/summary.txt ;;
    benchref) timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/benchref.json 2> gpurun_out/benchref.err
           echo "benchref exit $? : $(tail -c 400 gpurun_out/benchref.json)" | tee -a gpurun_out/summary.txt ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
              python bench.py --steps 1 --warmup 1 --depth 2 --no-cpu-baseline --no-gpu-baseline --no-graph > gpurun_out/ncu_list.log 2>&1
           echo "ncu_list exit $?" | tee -a gpurun_out/summary.txt ;;
    ncu_gemm) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 36 -c 12 -o gpurun_out/prof_gemm -f \
              python bench.py --steps 1 --warmup 1 --depth 1 --no-cpu-baseline --no-gpu-baseline --no-graph > gpurun_out/ncu_gemm.log 2>&1
           echo "ncu_gemm exit $?" | tee -a gpurun_out/summary.txt ;;
    ncu_short) timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_short -c 1 -o gpurun_out/prof_short -f \
              env KB_ONLY=elem python tools/kernel_bench.py > gpurun_out/ncu_short.log 2>&1
           echo "ncu_short exit $?" | tee -a gpurun_out/summary.txt ;;
    ncu_attn) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 2 -c 2 -o gpurun_out/prof_attn -f \
              python bench.py --steps 1 --warmup 1 --depth 1 --no-cpu-baseline --no-gpu-baseline --no-graph $NCU_BENCH_ARGS > gpurun_out/ncu_attn.log 2>&1
           echo "ncu_attn exit $?" | tee -a gpurun_out/summary.txt ;;
    *)     run "custom" 600 $g ;;
  esac
done
grep -h "parity\]\|PASS\|FAIL\|passed\|failed\|error\|watchdog" gpurun_out/*.log | tail -n 60
