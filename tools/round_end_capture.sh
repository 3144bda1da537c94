#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): GPU tests, bench (both arms), ncu launch
# list of one diffusion step and --set full captures of the two dominant kernels.
# Outputs go to gpurun_out/; the summaries are copied to profiles/ by hand afterwards.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_tests.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 2>> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench_reference.json
KREGEX='regex:gemm_bf16|attention_tcgen05|attention_combine|rmsnorm_film|sampler_step|step_advance'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 320 -c 137 --csv \
  --log-file gpurun_out/final_launches.csv python tools/profile_step.py > gpurun_out/final_prof_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05_kernel -s 25 -c 1 \
  -f -o gpurun_out/final_attn_cross python tools/profile_step.py > gpurun_out/final_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair -s 114 -c 1 \
  -f -o gpurun_out/final_gemm_wi python tools/profile_step.py > gpurun_out/final_ncu_gemm.log 2>&1
ls -la gpurun_out | tail -12
