#!/bin/bash
# A/B timing of two engine builds on the GPU box: tools/ab.sh NAME=path/to/lib.so ...   (fold-eval, score and Monte-Carlo rates)
export ADDAPT_B200_QUIET=1
for spec in "$@"; do
  name=${spec%%=*}; lib=${spec#*=}
  export ADDAPT_B200_LIB=$lib
  python tools/quick_time.py $name
  python bench.py --workload score --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$name', d['metric'], round(d['value']))"
  python bench.py --workload mc --batch 8192 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$name', d['metric'], round(d['value']))"
done
