# 8-GPU confirmation of the data-parallel bench (weak, then strong), each under its own timeout
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
S=$(date +%s)
timeout 240 $TR --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mg8_weak_graph.json 2> gpurun_out/r2_mg8_weak_graph.err; echo "weak graph exit $? after $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 240 $TR --master-port 29522 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/r2_mg8_strong_graph.json 2> gpurun_out/r2_mg8_strong_graph.err; echo "strong graph exit $? after $(( $(date +%s) - S )) s"
for f in gpurun_out/r2_mg8_*.json; do echo $f; grep '^{' $f | cut -c1-230; done
grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/r2_mg8_weak_graph.err | tail -5
