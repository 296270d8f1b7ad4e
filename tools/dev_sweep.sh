run() { timeout 120 python bench.py --steps 10 --warmup 3 --no-match --no-cpu --no-bow 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
echo -n "normal: "; run
echo -n "skip h2d: "; ORBX_DEBUG_SKIP_H2D=1 run
echo -n "skip d2h: "; ORBX_DEBUG_SKIP_D2H=1 run
echo -n "skip both: "; ORBX_DEBUG_SKIP_H2D=1 ORBX_DEBUG_SKIP_D2H=1 run
echo -n "skip both, plan 256x4: "; ORBX_CHUNK_PLAN=256,256,256,256 ORBX_DEBUG_SKIP_H2D=1 ORBX_DEBUG_SKIP_D2H=1 run
