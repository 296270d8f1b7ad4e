run() { timeout 200 python bench.py --workload $1 --batch $2 --steps 8 --warmup 3 --no-match --no-cpu --no-bow 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
for ns in 1 2 3 4; do echo -n "tum1 B=1024 nsub=$ns: "; ORBX_NSUB=$ns run tum1 1024; done
for ns in 1 2; do echo -n "tum1 B=2048 nsub=$ns: "; ORBX_NSUB=$ns run tum1 2048; done
