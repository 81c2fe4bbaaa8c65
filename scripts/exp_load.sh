for l in 1.3 1.6 2.0 2.5 3.2; do echo "load=$l"; GSB_PROB_SLOTS=8 GSB_PROB_LOAD=$l timeout 300 python bench.py --steps 2 --warmup 3 --batch 32 --cpu-sample 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['kernel_ms'], d['roofline']['avg_launch_ms'])"; done
