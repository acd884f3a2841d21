#!/bin/bash
# Development aid: multi-GPU session (gpurun --gpus N): parity of the sharded run, weak-scaling bench with self-check, c4 strong scaling.
n=${1:-2}
tag=${2:-mg}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/multi_gpu_check.py > gpurun_out/${tag}_check_n$n.log 2>&1; echo "multi_gpu_check rc=$?"; grep "rank " gpurun_out/${tag}_check_n$n.log | head -8
CMLBA_NCCL_ONLY=1 timeout 300 $TR --master-port 29534 tools/multi_gpu_check.py > gpurun_out/${tag}_check_nccl_n$n.log 2>&1; echo "multi_gpu_check (nccl) rc=$?"
timeout 300 $TR --master-port 29535 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/${tag}_bench_weak_n$n.json 2> gpurun_out/${tag}_bench_weak_n$n.err; echo "bench weak rc=$?"
timeout 300 $TR --master-port 29536 bench.py --gpus $n --steps 30 --warmup 5 --workload c4 --scaling strong > gpurun_out/${tag}_bench_c4strong_n$n.json 2> gpurun_out/${tag}_bench_c4strong_n$n.err; echo "bench c4 strong rc=$?"
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --workload c4 --scaling strong --no-cpu-baseline > gpurun_out/${tag}_bench_c4strong_n1.json 2> gpurun_out/${tag}_bench_c4strong_n1.err; echo "bench c4 n1 rc=$?"
if [ "$n" == "2" ]; then timeout 600 $TR --master-port 29537 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_n$n.json 2> gpurun_out/${tag}_bench_ref_n$n.err; echo "reference arm under torchrun rc=$? lines=$(wc -l < gpurun_out/${tag}_bench_ref_n$n.json)"; fi
for f in gpurun_out/${tag}_bench_weak_n$n.json gpurun_out/${tag}_bench_c4strong_n$n.json gpurun_out/${tag}_bench_c4strong_n1.json; do python -c "
import json
try:
    d=json.loads([l for l in open('$f').read().splitlines() if l.startswith('{')][-1]); print('$f', 'value %.3e' % d['value'], 'ms %.1f us' % (d['ms_per_step']*1e3), 'e2e %.3e' % d['e2e']['value'], d.get('multi_gpu_check'))
except Exception as e: print('$f ERR', open('$f'.replace('.json','.err')).read()[-600:])
"; done
