#!/bin/bash
# Shorter A/B than tools/ab.sh: fold-eval / inside / MFE rates per build, in the order given (repeat a name to interleave);
# tools/ab_quick.sh NAME=path/to/lib.so ...
export ADDAPT_B200_QUIET=1
for spec in "$@"; do
  name=${spec%%=*}; lib=${spec#*=}
  ADDAPT_B200_LIB=$lib python tools/quick_time.py $name
done
