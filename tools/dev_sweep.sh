run() { timeout 200 python bench.py --workload $1 --batch $2 --steps 8 --warmup 3 --no-match --no-cpu --no-bow 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
for cfg in "tum1 1024" "tum1 256" "tum1 128" "kitti 256" "euroc 512" "4k 64"; do set -- $cfg; echo -n "$1 B=$2 auto: "; run $1 $2; done
for cfg in "tum1 256" "tum1 128"; do set -- $cfg; echo -n "$1 B=$2 nsub=4: "; ORBX_NSUB=4 run $1 $2; done
