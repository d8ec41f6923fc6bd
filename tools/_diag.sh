mkdir -p gpurun_out/s3
python tools/power_probe.py > gpurun_out/s3/power_probe.log 2>&1
for d in 0 1 2 3; do DASR_LIB_PATH=depth_aware_endoscopy_sr_b200/libdasr_b200_prof.so DASR_DBG=$d python tools/prof_sean.py; done > gpurun_out/s3/prof_sean_dbg.log 2>&1
grep -v Warn gpurun_out/s3/power_probe.log | tail -40; cat gpurun_out/s3/prof_sean_dbg.log
