mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_mg_test.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_mg_test.log
tail -4 gpurun_out/r2_mg_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mg2_weak_graph.json 2> gpurun_out/r2_mg2_weak_graph.err; echo "weak graph exit $?"
HRV_MULTI_GRAPH=0 timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mg2_weak_eager.json 2> gpurun_out/r2_mg2_weak_eager.err; echo "weak eager exit $?"
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/r2_mg2_strong_graph.json 2> gpurun_out/r2_mg2_strong_graph.err; echo "strong graph exit $?"
for f in gpurun_out/r2_mg2_*.json; do echo $f; cut -c1-260 $f; done
tail -5 gpurun_out/r2_mg2_weak_graph.err
