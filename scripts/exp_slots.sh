for s in 1 2 4 8; do echo "slots=$s"; GSB_PROB_SLOTS=$s timeout 300 python bench.py --steps 2 --warmup 3 --batch 32 --cpu-sample 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['kernel_ms'], d['roofline']['avg_launch_ms'])"; done
