#!/bin/bash
# round-2 `ncu --set full` captures of the kernels the north_star names (HBM-bound assembly / solve kernels at the C4
# shape, G = 1.07 GB >> L2; the Cholesky trailing updates; the batched kernel).  Run on the GPU box: bash tools/run_ncu_r2.sh
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:gemv_n_kernel -s 2 -c 1 -o gpurun_out/r2_gemv_n_c4 python tools/gpu_kernels.py 7 8192 16384 > gpurun_out/r2_ncu_gemv_n.log 2>&1
$N -k regex:gemv_t_kernel -s 2 -c 1 -o gpurun_out/r2_gemv_t_c4 python tools/gpu_kernels.py 8 8192 16384 > gpurun_out/r2_ncu_gemv_t.log 2>&1
$N -k regex:ruiz_fused_kernel -s 1 -c 1 -o gpurun_out/r2_ruiz_n8192 python tools/gpu_kernels.py 6 8192 0 > gpurun_out/r2_ncu_ruiz.log 2>&1
$N -k regex:trsv_fwd_wave -s 1 -c 1 -o gpurun_out/r2_trsv_fwd_n8192 python tools/gpu_kernels.py 5 8192 0 > gpurun_out/r2_ncu_trsv_fwd.log 2>&1
$N -k regex:trsv_bwd_wave -s 1 -c 1 -o gpurun_out/r2_trsv_bwd_n8192 python tools/gpu_kernels.py 5 8192 0 > gpurun_out/r2_ncu_trsv_bwd.log 2>&1
$N -k regex:gemm_dmma_streamk_kernel -s 1 -c 1 -o gpurun_out/r2_chol_trailing_4096_k4096 python tools/gpu_kernels.py 2 4096 4096 > gpurun_out/r2_ncu_trail.log 2>&1
$N -k regex:gemm_dmma_kernel -s 1 -c 1 -o gpurun_out/r2_chol_rank128_8064 python tools/gpu_kernels.py 2 8064 128 > gpurun_out/r2_ncu_rank128.log 2>&1
$N -k regex:batched_barrier_kernel -s 1 -c 1 -o gpurun_out/r2_batched_1024 python tools/gpu_batch.py 1024 > gpurun_out/r2_ncu_batched.log 2>&1
$N -k regex:pd_linesearch_kernel -s 1 -c 1 -o gpurun_out/r2_pd_linesearch_c4 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-batched --no-legs --no-e2e > gpurun_out/r2_ncu_ls.log 2>&1
ls -la gpurun_out/*.ncu-rep
