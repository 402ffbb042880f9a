# A/B of the kernel-selection rules on one GPU: thin-tile probes and whole-step bench lines
mkdir -p gpurun_out
C=12,13,14,15,16,17,3
HRV_CONV_PAIR_MINBN=144 HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab_probe_r2a.txt 2>&1     # rules of the previous commit
HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab_probe_default.txt 2>&1
HRV_CONV_PAIR_MINBN=32 HRV_CONV_PAIR_OVER_PIXN_MAXBN=64 HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab_probe_min32.txt 2>&1
HRV_CONV_PAIR_MINBN=32 HRV_CONV_PAIR_OVER_PIXN_MAXBN=128 HRV_CONV_PIXN=1 HRV_PROBE_CASES=$C,9,10,11 python tools/conv_stall_probe.py 8 > gpurun_out/r2_ab_probe_over128.txt 2>&1
python tools/pair_smoke.py > gpurun_out/r2_ab_pair_smoke.txt 2>&1; tail -1 gpurun_out/r2_ab_pair_smoke.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2_ab_prof_default.csv > gpurun_out/r2_ab_bench_default.json 2> gpurun_out/r2_ab_bench_default.err
HRV_CONV_PAIR_MINBN=32 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2_ab_prof_min32.csv > gpurun_out/r2_ab_bench_min32.json 2> gpurun_out/r2_ab_bench_min32.err
HRV_CONV_PAIR_MINBN=32 HRV_CONV_PAIR_OVER_PIXN_MAXBN=128 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2_ab_prof_over128.csv > gpurun_out/r2_ab_bench_over128.json 2> gpurun_out/r2_ab_bench_over128.err
for f in gpurun_out/r2_ab_bench_*.json; do echo $f; grep '^{' $f | cut -c1-190; done
for f in gpurun_out/r2_ab_probe_*.txt; do echo $f; cat $f; done
