# owner-mode epilogue of the one-CTA kernel + feature-matching kernel path: correctness, then A/B inside the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_t9.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_t9.log
C=13,14,17,12
HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C HRV_CONV_EPI_OWN=0 python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab2_probe_own0.txt 2>&1
HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab2_probe_own1.txt 2>&1
HRV_CONV_EPI_OWN=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2_ab2_prof_own0.csv > gpurun_out/r2_ab2_bench_own0.json 2> gpurun_out/r2_ab2_bench_own0.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2_ab2_prof_own1.csv > gpurun_out/r2_ab2_bench_own1.json 2> gpurun_out/r2_ab2_bench_own1.err
tail -2 gpurun_out/r2_ab2_bench_own1.err
for f in gpurun_out/r2_ab2_bench_*.json; do echo $f; grep '^{' $f | cut -c1-190; done
cat gpurun_out/r2_ab2_probe_own0.txt gpurun_out/r2_ab2_probe_own1.txt
