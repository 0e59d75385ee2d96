#!/bin/bash
# usage: scripts/probe_variants.sh v1 v2 ...   (times the 16-B and 24-B stage of cfg3 for each library variant;
# environment such as FRB_MARCH_PREFETCH / FRB_MARCH_PFDIST is passed through)
for v in "$@"; do
  L=fluxreconstruction.jl_b200/lib/variants/libfrb200_$v.so
  echo "== $v pf=${FRB_MARCH_PREFETCH:-1} dist=${FRB_MARCH_PFDIST:-0} $(FRB200_LIB=$L python scripts/probe_cfg3.py 2048 march 2>&1 | grep stage_kind | awk '{printf "kind%s %s ms  ", substr($2,12), $3}')"
done
