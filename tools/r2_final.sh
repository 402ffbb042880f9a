# final validation of the round: whole GPU suite, smoke, default bench line (with cpu_baseline), the other workloads, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/r2_t10.log 2>&1; echo "pytest exit $?"; grep -E "^FAILED|^ERROR|^SKIPPED|passed|failed" gpurun_out/r2_t10.log | tail -12
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; grep smoke: gpurun_out/r2_smoke.log | cut -c1-300
S=$(date +%s); python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $? after $(( $(date +%s) - S )) s"; grep '^{' gpurun_out/r2_bench_final.json | cut -c1-200
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err; echo "reference exit $? after $(( $(date +%s) - S )) s"; cut -c1-160 gpurun_out/r2_bench_reference_final.json
for wl in gen_fwd train_stage1 pipeline; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final_$wl.json 2> gpurun_out/r2_bench_final_$wl.err; echo "$wl exit $?"; grep '^{' gpurun_out/r2_bench_final_$wl.json | cut -c1-170
done
python tools/hbm_bench.py 8 10 > gpurun_out/r2_hbm_bench_final.txt 2>&1
