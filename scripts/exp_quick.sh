# usage: exp_quick.sh "ENV=VAL ..." ...   one bench line per configuration
for cfg in "$@"; do echo "cfg: $cfg"; env $cfg timeout 300 python bench.py --steps 2 --warmup 3 --batch 32 --cpu-sample 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,2) for k,v in d['kernel_ms'].items()}, d['roofline']['avg_launch_ms'], d['retries'])"; done
