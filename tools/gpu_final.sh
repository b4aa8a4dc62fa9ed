# Round-end validation on one B200 (run through gpurun): whole -m gpu suite, the tf32 GEMM microbenchmark, the train bench and
# smoke(); optionally the same tf32 / train steps against an alternative build (ALT=path/to/libdinounet_b200_x.so, see
# csrc/build.py's B2U_OUT_SUFFIX).  Everything small goes to gpurun_out/ (never write ncu reports there: 64 MiB cap).
mkdir -p gpurun_out
rm -f gpurun_out/f_rc.txt
timeout 480 python -m pytest tests -q -m gpu -x > gpurun_out/f_tests.log 2>&1; echo "rc suite $?" >> gpurun_out/f_rc.txt
timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/f_micro_default.log 2>&1; echo "rc micro $?" >> gpurun_out/f_rc.txt
timeout 200 python bench.py --mode train --model dinounet_b --batch 64 --steps 4 --warmup 3 > gpurun_out/f_train_default.json 2> gpurun_out/f_train_default.err; echo "rc train $?" >> gpurun_out/f_rc.txt
if [ -n "$ALT" ] && [ -f "$ALT" ]; then
  DINOUNET_B200_LIB=$ALT timeout 100 python tools/bench_tf32_gemm.py 8 > gpurun_out/f_micro_alt.log 2>&1; echo "rc micro alt $?" >> gpurun_out/f_rc.txt
  DINOUNET_B200_LIB=$ALT timeout 200 python bench.py --mode train --model dinounet_b --batch 64 --steps 4 --warmup 3 > gpurun_out/f_train_alt.json 2> gpurun_out/f_train_alt.err; echo "rc train alt $?" >> gpurun_out/f_rc.txt
  DINOUNET_B200_LIB=$ALT timeout 200 python -m pytest tests/test_gpu_tf32_gemm.py tests/test_gpu_train.py -q > gpurun_out/f_tests_alt.log 2>&1; echo "rc tests alt $?" >> gpurun_out/f_rc.txt
fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "rc smoke $?" >> gpurun_out/f_rc.txt
tail -4 gpurun_out/f_tests.log; cat gpurun_out/f_rc.txt; du -sh gpurun_out
