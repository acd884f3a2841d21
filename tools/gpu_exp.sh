#!/bin/bash
# Development aid: experiment session on one GPU box (parity tests, bench under development switches, traces, kernel timeline).
#   gpurun --timeout 1200 -- 'bash tools/gpu_exp.sh <tag>'
tag=${1:-exp}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${tag}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity_report.jsonl 2>/dev/null
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_$name.json'));k=d['kernel_ms'];print('$name', 'pass', round(d['ms_per_step']*1e3,1), 'warm', round(d['ms_per_step_l2_warm']*1e3,1), {a:round(b*1e3,1) for a,b in k.items() if a!='note'}, d['run'], 'e2e ms', round(d['e2e']['ms_per_step'],3))"; }
b default X=1
b nointerleave CMLBA_NO_INTERLEAVE=1
b nofork CMLBA_NO_FORK=1
LT_COLD=0 CMLBA_LT_MODE=2 timeout 120 python tools/lt_trace.py > gpurun_out/${tag}_trace_warm_m2.txt 2>&1
echo "== warm trace"; grep -v "Warn\|_ureduce\|nanm" gpurun_out/${tag}_trace_warm_m2.txt | head -8
LT_COLD=1 CMLBA_LT_MODE=2 timeout 120 python tools/lt_trace.py > gpurun_out/${tag}_trace_cold_m2.txt 2>&1
echo "== cold trace"; grep -v "Warn\|_ureduce\|nanm" gpurun_out/${tag}_trace_cold_m2.txt | head -8
