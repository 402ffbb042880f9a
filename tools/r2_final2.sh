# last validation of the round on the final commit: whole GPU suite, smoke, default bench line, inference lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/r2_t12.log 2>&1; echo "pytest exit $?"; grep -E "^FAILED|^ERROR|^SKIPPED|passed|failed" gpurun_out/r2_t12.log | tail -8
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke2.log 2>&1; echo "smoke exit $?"; grep smoke: gpurun_out/r2_smoke2.log | cut -c1-300
python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; echo "bench exit $?"; grep '^{' gpurun_out/r2_bench_final2.json | cut -c1-200
for wl in gen_fwd pipeline train_stage1; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final2_$wl.json 2> gpurun_out/r2_bench_final2_$wl.err; echo "$wl exit $?"; grep '^{' gpurun_out/r2_bench_final2_$wl.json | cut -c1-170
done
