# final validation of a round: whole -m gpu suite on the default build, the tf32 tier's A/B against the row-mode-only build,
# train bench on both, smoke().  Everything small goes to gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/f_rc.txt
timeout 480 python -m pytest tests -q -m gpu -x > gpurun_out/f_tests.log 2>&1; echo "rc suite $?" >> gpurun_out/f_rc.txt
timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/f_micro_default.log 2>&1; echo "rc micro $?" >> gpurun_out/f_rc.txt
ALT=$PWD/dinounet_b200/libdinounet_b200_pf12.so
DINOUNET_B200_LIB=$ALT timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/f_micro_pf12.log 2>&1; echo "rc micro12 $?" >> gpurun_out/f_rc.txt
timeout 200 python bench.py --mode train --model dinounet_b --batch 64 --steps 4 --warmup 3 > gpurun_out/f_train_default.json 2> gpurun_out/f_train_default.err; echo "rc train $?" >> gpurun_out/f_rc.txt
DINOUNET_B200_LIB=$ALT timeout 200 python bench.py --mode train --model dinounet_b --batch 64 --steps 4 --warmup 3 > gpurun_out/f_train_pf12.json 2> gpurun_out/f_train_pf12.err; echo "rc train12 $?" >> gpurun_out/f_rc.txt
DINOUNET_B200_LIB=$ALT timeout 200 python -m pytest tests/test_gpu_tf32_gemm.py tests/test_gpu_train.py -q > gpurun_out/f_tests_pf12.log 2>&1; echo "rc tests12 $?" >> gpurun_out/f_rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "rc smoke $?" >> gpurun_out/f_rc.txt
tail -4 gpurun_out/f_tests.log; cat gpurun_out/f_rc.txt; du -sh gpurun_out
