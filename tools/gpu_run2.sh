mkdir -p gpurun_out
for v in pf12 pf22 pf32; do
  DINOUNET_B200_LIB=$PWD/dinounet_b200/libdinounet_b200_$v.so timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/r2_micro_$v.log 2>&1; echo "rc $v $?" >> gpurun_out/r2_rc.txt
done
timeout 150 python tools/grad_tier_report.py > gpurun_out/r2_grad_report.log 2>&1; echo "rc report $?" >> gpurun_out/r2_rc.txt
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_one_step.py dinounet_b 32 > gpurun_out/r2_ncu_train.log 2>&1; echo "rc ncutrain $?" >> gpurun_out/r2_rc.txt
cat gpurun_out/r2_rc.txt; du -sh gpurun_out
