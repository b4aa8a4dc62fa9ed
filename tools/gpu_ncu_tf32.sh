# ncu --set full of three tf32 GEMM launches (report kept in /tmp on the box; only the exported pages + a small .ncu-rep come back)
mkdir -p gpurun_out
rm -f gpurun_out/h_rc.txt
timeout 85 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 3 -c 3 -o /tmp/tf32_full python tools/ncu_tf32_three.py > gpurun_out/h_ncu_full.log 2>&1; echo "rc ncufull $?" >> gpurun_out/h_rc.txt
if [ -f /tmp/tf32_full.ncu-rep ]; then
  ncu -i /tmp/tf32_full.ncu-rep --page raw --csv > gpurun_out/h_tf32_full_raw.csv 2>/dev/null
  ncu -i /tmp/tf32_full.ncu-rep --page details > gpurun_out/h_tf32_full_details.txt 2>/dev/null
  sz=$(stat -c %s /tmp/tf32_full.ncu-rep); echo "rep bytes $sz" >> gpurun_out/h_rc.txt
  if [ "$sz" -lt 25000000 ]; then cp /tmp/tf32_full.ncu-rep gpurun_out/h_tf32_full.ncu-rep; fi
fi
cat gpurun_out/h_rc.txt; du -sh gpurun_out
