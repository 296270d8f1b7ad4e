# end-of-change check on the GPU box: full GPU suite, default bench line, the other BASELINE shapes
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
for w in euroc kitti 4k; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-match --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms_per_batch'])"
done
