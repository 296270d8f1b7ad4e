# end-of-change check on the GPU box: GPU suite (or a subset via $1), bench line, the other BASELINE shapes
timeout 900 python -m pytest ${1:-tests} -m gpu -x -q 2>&1 | tail -4
for w in tum1 euroc kitti 4k; do
  timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-match --no-cpu --no-bow > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline']['stage_ms_per_batch'].items()})"
done
