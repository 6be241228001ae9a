#!/bin/bash
# JIT CTA-size / register-cap sweep.  Usage: tools/sweep_jit.sh scene n "T:MINB T:MINB ..." [pts-list]
SCENE=${1:-cfg_planetary}; N=${2:-512}; COMBOS=${3:-"128:0 128:8 256:4 512:2 512:0 256:2"}
export PROBE_INTERP=0 PROBE_JIT_PTS=${4:-1,2,4}
for C in $COMBOS; do
  T=${C%%:*}; MB=${C##*:}
  if [ "$MB" != "0" ]; then export CODECAD_B200_JIT_MINB=$MB; else unset CODECAD_B200_JIT_MINB; fi
  echo "== threads $T minb $MB"
  CODECAD_B200_JIT_THREADS=$T timeout 600 python tools/gpu_probe.py $SCENE $N 2>&1 | grep -v "^device\|micro-ops"
done
