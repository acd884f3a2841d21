for m in 0 255 60 56 48 32 124 188; do
CMLBA_PDL_MASK=$m timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/pdl_$m.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/pdl_$m.json'));k=d['kernel_ms'];print('mask $m pass', round(d['ms_per_step']*1e3,1), 'warm', round(d['ms_per_step_l2_warm']*1e3,1), 'lin', round(k['linearize_accumulate']*1e3,1), 'schur', round(k['schur']*1e3,1), 'stitch', round(k['stitch_assemble']*1e3,1), 'run_ms', round(d['run']['gpu_ms'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],3))"
done
