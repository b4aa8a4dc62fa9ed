mkdir -p gpurun_out
rm -f gpurun_out/g_rc.txt
ALT=$PWD/dinounet_b200/libdinounet_b200_blkrt.so
timeout 200 python -m pytest tests/test_gpu_tf32_gemm.py tests/test_gpu_train.py -q > gpurun_out/g_tests.log 2>&1; echo "rc tests $?" >> gpurun_out/g_rc.txt
timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/g_micro_default.log 2>&1; echo "rc micro $?" >> gpurun_out/g_rc.txt
DINOUNET_B200_LIB=$ALT timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/g_micro_blkrt.log 2>&1; echo "rc microalt $?" >> gpurun_out/g_rc.txt
timeout 200 python bench.py --mode train --model dinounet_b --batch 64 --steps 4 --warmup 3 > gpurun_out/g_train_default.json 2> gpurun_out/g_train_default.err; echo "rc train $?" >> gpurun_out/g_rc.txt
DINOUNET_B200_LIB=$ALT timeout 200 python bench.py --mode train --model dinounet_b --batch 64 --steps 4 --warmup 3 > gpurun_out/g_train_blkrt.json 2> gpurun_out/g_train_blkrt.err; echo "rc trainalt $?" >> gpurun_out/g_rc.txt
tail -3 gpurun_out/g_tests.log; cat gpurun_out/g_rc.txt; du -sh gpurun_out
