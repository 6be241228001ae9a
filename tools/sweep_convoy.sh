#!/bin/bash
# JIT CTA-size / convoy-barrier sweep (planetary 512^3).  Usage: tools/sweep_convoy.sh [scene] [n]
SCENE=${1:-cfg_planetary}; N=${2:-512}
export PROBE_INTERP=0 PROBE_JIT_PTS=${PROBE_JIT_PTS:-1,2}
for T in 128 256 512 1024; do
  for S in 0 1 4; do
    for MB in "" 2; do
      if [ -n "$MB" ]; then
        # min CTAs per SM so that the register cap is 64 (only meaningful combos)
        case $T in 128) MINB=8;; 256) MINB=4;; 512) MINB=2;; 1024) continue;; esac
        export CODECAD_B200_JIT_MINB=$MINB
      else
        unset CODECAD_B200_JIT_MINB
      fi
      echo "== threads $T sync $S minb ${CODECAD_B200_JIT_MINB:-none}"
      CODECAD_B200_JIT_THREADS=$T CODECAD_B200_JIT_SYNC=$S timeout 300 python tools/gpu_probe.py $SCENE $N 2>&1 | grep -v "^device\|micro-ops"
    done
  done
done
