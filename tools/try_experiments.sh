#!/bin/bash
# Builds the experiment variants prepared at the end of round 1 (-DADP_CORNER_TABLE, -DADP_QM_SCAN, both,
# and -DADP_DESIGN_ONE_CST for the score / Monte-Carlo kernels)
# and prints the one GPU command that checks their parity and times them against the product build.
set -e
cd "$(dirname "$0")/.."
tools/variant_build.sh ctab -DADP_CORNER_TABLE &
tools/variant_build.sh qs -DADP_QM_SCAN &
tools/variant_build.sh ctabqs -DADP_CORNER_TABLE -DADP_QM_SCAN &
tools/variant_build.sh onecst -DADP_DESIGN_ONE_CST &
tools/variant_build.sh gcap3 -DADP_CELL_GROUP_CAP_LG=3 &
tools/variant_build.sh gcap4 -DADP_CELL_GROUP_CAP_LG=4 &
wait
rm -f build/*/*.o
cat <<'CMD'
gpurun --timeout 900 -- 'python tools/qt2.py base; python tools/qt_mfe.py base; for v in ctab qs ctabqs gcap3 gcap4; do export ADDAPT_B200_LIB=build/$v/libaddapt_b200_$v.so; python -m pytest tests -m gpu -x -q | tail -1; python tools/qt2.py $v; python tools/qt_mfe.py $v; done; unset ADDAPT_B200_LIB; for v in base onecst; do [ $v = base ] || export ADDAPT_B200_LIB=build/$v/libaddapt_b200_$v.so; python bench.py --workload score --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-90; python bench.py --workload mc --batch 8192 --steps 1 --warmup 2 --no-cpu-baseline | cut -c1-90; done'
CMD
