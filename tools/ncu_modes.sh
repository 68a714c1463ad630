#!/bin/bash
# cycles / duration / instruction count of K1 under precision modes and debug masks (ncu, no clock control)
# usage: bash tools/ncu_modes.sh "0:0 0:64 1:0"   (mode:dbg pairs)
for md in ${1:-0:0 0:2 0:4 1:0 1:2 1:4}; do mode=${md%%:*}; dbg=${md##*:}
  NPLDA_TC_MODE=$mode NPLDA_TC_DEBUG=$dbg ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:score_tc -s 3 -c 2 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --kernel tc 2>/dev/null | grep -E "score_tc" | awk -F'","' -v m=$mode -v d=$dbg '{printf "mode %s dbg %s %s %s\n", m, d, $(NF-2), $NF}' | tr -d '"' | paste -sd' '
done
