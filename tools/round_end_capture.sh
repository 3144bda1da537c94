#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): ncu launch list of one diffusion step and
# --set full captures of the dominant kernels.  Outputs go to gpurun_out/; the summaries are copied
# to profiles/ afterwards (tools/summarise_ncu.py).
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2}
KREGEX='regex:gemm_bf16|attention_tcgen05|attention_combine|rmsnorm_film|sampler_step'
# encode = 109 GEMM + 24 attention + 50 norm launches, then one warm-up step of 101 kernels
# (deferred normalisation: 74 GEMM + 24 attention + 2 norm + sampler; 136 with MSD_FUSED_NORM=0)
STEP_KERNELS=${STEP_KERNELS:-101}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s $((183 + STEP_KERNELS)) -c $STEP_KERNELS --csv \
  --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py > gpurun_out/${TAG}_prof_step.log 2>&1
# cross-attention of layer 0 (128-key instance with the long/short split), self-attention (64-key
# instance, two CTAs per SM), wi GEMM (gated GELU), self-out GEMM (TMA reduce-add epilogue)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05_kernel -s 25 -c 1 \
  -f -o gpurun_out/${TAG}_attn_cross python tools/profile_step.py > gpurun_out/${TAG}_ncu_attn_cross.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05_kernel -s 24 -c 1 \
  -f -o gpurun_out/${TAG}_attn_self python tools/profile_step.py > gpurun_out/${TAG}_ncu_attn_self.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair -s 114 -c 1 \
  -f -o gpurun_out/${TAG}_gemm_wi python tools/profile_step.py > gpurun_out/${TAG}_ncu_gemm_wi.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair -s 111 -c 1 \
  -f -o gpurun_out/${TAG}_gemm_out python tools/profile_step.py > gpurun_out/${TAG}_ncu_gemm_out.log 2>&1
ls -la gpurun_out | tail -12
