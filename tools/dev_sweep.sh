run() { timeout 200 python bench.py --workload $1 --steps 8 --warmup 3 --no-match --no-cpu --no-bow 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
for w in kitti euroc 4k; do for ns in 8 5 3 2; do echo -n "$w nsteady=$ns: "; ORBX_NSTEADY=$ns run $w; done; done
