for cfg in "2 4 8" "3 4 8" "4 4 8" "3 6 12" "4 8 12" "3 3 6" "2 4 16" "4 8 16" "2 8 8"; do
  set -- $cfg
  echo -n "nside=$1 nsub=$2 nsteady=$3: "
  ORBX_NSIDE=$1 ORBX_NSUB=$2 ORBX_NSTEADY=$3 timeout 120 python bench.py --steps 10 --warmup 3 --no-match --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"
done
