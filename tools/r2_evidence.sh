# Round-2 evidence pass (one GPU): stall probes, ncu captures, launch list + DRAM traffic of one step, the other workloads, the context arms
mkdir -p gpurun_out
HRV_PROBE_CASES=9,10,11,12,3 python tools/conv_stall_probe.py 8 > gpurun_out/r2_stall_bn128_single.txt 2>&1
HRV_CONV_PAIR_MINBN=64 HRV_PROBE_CASES=9,10,11,12,3 python tools/conv_stall_probe.py 8 > gpurun_out/r2_stall_bn128_pair.txt 2>&1
python tools/conv_stall_probe.py 8 > gpurun_out/r2_conv_stall_final_b8.txt 2>&1
HRV_PROBE_ITERS=0 HRV_PROBE_CASES=0,1,2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_pair -c 3 -f -o gpurun_out/r2_ncu_pair python tools/conv_stall_probe.py 8 > gpurun_out/r2_ncu_pair.log 2>&1
ncu -i gpurun_out/r2_ncu_pair.ncu-rep --page raw --csv > gpurun_out/r2_ncu_pair_raw.csv 2>/dev/null
timeout 900 ncu --nvtx --nvtx-include "hrv_profile_step" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/r2_launches.csv)"
python tools/hbm_bench.py 8 10 > gpurun_out/r2_hbm_bench_b8.txt 2>&1
for wl in gen_fwd train_stage1 pipeline; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "$wl exit $?"; cut -c1-200 gpurun_out/r2_bench_$wl.json
done
timeout 300 python bench.py --impl torch_gpu --steps 3 --warmup 2 > gpurun_out/r2_bench_torch_gpu.json 2> gpurun_out/r2_bench_torch_gpu.err; echo "torch_gpu exit $?"; cut -c1-250 gpurun_out/r2_bench_torch_gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference exit $?"; cut -c1-400 gpurun_out/r2_bench_reference.json
cat gpurun_out/r2_stall_bn128_single.txt gpurun_out/r2_stall_bn128_pair.txt
