#!/bin/bash
# Development aid: one GPU-box session = parity tests + bench line + ncu launch list + ncu full capture of the sampling kernel.
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh <tag> [quick|variants]'
# Everything lands in gpurun_out/<tag>_*; every step runs under its own timeout so that a hung kernel cannot hold the box.
tag=${1:-dev}
mode=${2:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 120 python tools/gpu_check.py tiny > gpurun_out/${tag}_check_tiny.log 2>&1
echo "gpu_check rc=$?" >> gpurun_out/${tag}_check_tiny.log
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity_report.jsonl 2>/dev/null
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
if [ "$mode" == "variants" ]; then
  for v in 0 2; do
    for nt in 0; do
      if [ $nt == 1 ]; then export CMLBA_NO_TMA=1; else unset CMLBA_NO_TMA; fi
      CMLBA_LT_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${tag}_bench_v${v}_$nt.json 2> gpurun_out/${tag}_bench_v${v}_$nt.err
      python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_v${v}_$nt.json'));print('variant $v no_tma=$nt', round(d['ms_per_step']*1e3,1), 'lin', round(d['kernel_ms']['linearize']*1e3,1), 'warm', round(d['kernel_ms']['linearize_l2_warm']*1e3,1), d['run'], round(d['e2e']['ms_per_step'],3))"
    done
  done
  unset CMLBA_NO_TMA
  LT_COLD=0 CMLBA_LT_MODE=2 timeout 120 python tools/lt_trace.py > gpurun_out/${tag}_trace_warm.txt 2>&1
  LT_COLD=1 CMLBA_LT_MODE=2 timeout 120 python tools/lt_trace.py > gpurun_out/${tag}_trace_cold.txt 2>&1
  CMLBA_LT_MODE=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${tag}_bench_stream.json 2> gpurun_out/${tag}_bench_stream.err
  python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_stream.json'));print('stream-only', d['ms_per_step'], d['kernel_ms'])"
  CMLBA_LT_EXACT=1 timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
fi
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json | head -c 3000
if [ "$mode" != "quick" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_l.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'linearize_tile|accumulate|schur|stitch|assemble|solve|post_lin|point_step|bin_' -s 24 -c 16 -o gpurun_out/${tag}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_f.log 2>&1
  ls -la gpurun_out/${tag}_* | tail -20
fi
