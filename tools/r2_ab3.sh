# unrolled tap-group MMA issue: correctness, then A/B (probe + whole step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_t11.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_t11.log
C=0,1,2,3,13,14,15,12,5
HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C HRV_CONV_TAP_GROUP=0 python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab3_probe_grp0.txt 2>&1
HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab3_probe_grp1.txt 2>&1
HRV_CONV_TAP_GROUP=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab3_bench_grp0.json 2> gpurun_out/r2_ab3_bench_grp0.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2_ab3_prof_grp1.csv > gpurun_out/r2_ab3_bench_grp1.json 2> gpurun_out/r2_ab3_bench_grp1.err
tail -2 gpurun_out/r2_ab3_bench_grp1.err
for f in gpurun_out/r2_ab3_bench_*.json; do echo $f; grep '^{' $f | cut -c1-190; done
cat gpurun_out/r2_ab3_probe_grp0.txt gpurun_out/r2_ab3_probe_grp1.txt
