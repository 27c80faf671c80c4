#!/bin/bash
# north_star: an ncu capture (--set full) for every hand-written kernel of the hot path; summaries -> profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu
NCU="ncu --profile-from-start off --set full --clock-control none -f"
run() { # name regex count model
  timeout -k 5 400 $NCU -k "regex:$2" -c $3 -o /tmp/r02_$1 python tools/profile_step.py $4 > gpurun_out/ncu/r02_$1.log 2>&1; echo "$1 rc=$?"
  # only the text summary travels back (gpurun_out is capped at 64 MiB; a --set full report is 1-2 MB per launch)
  python tools/ncu_summary.py /tmp/r02_$1.ncu-rep gpurun_out/ncu/r02_ncu_$1.txt > /dev/null 2>&1; rm -f /tmp/r02_$1.ncu-rep
}
run deepfm_step "fm_embed|segsum|radix|heads_|long_list|epoch_rows|tc_gemm|fc_dz|fc1_|logit_loss|dropout_mask|opt_dense_grad|splitk|epoch_tick" 60 deepfm
run dcn_cross "cross_" 4 dcn
run din_kernels "din_pool|gather_scale|bag_sum|group_sum|scale_rows|axpby" 24 din
run pnn_kernels "pnn_" 4 pnn
run afm_kernels "afm_" 8 afm
run bn_kernels "bn_" 8 deepfm_bn
run libsvm_kernels "ls_" 6 libsvm
run wide_deep_kernels "wd_" 6 wide_deep
ls -la gpurun_out/ncu | head -30
