#!/bin/bash
V=fluxreconstruction.jl_b200/lib/variants
for v in default nsold e6 e8 default nsold; do
  echo "== ns $v"; if [ $v = default ]; then python scripts/ns_probe.py; else FRB200_LIB=$V/libfrb200_$v.so python scripts/ns_probe.py; fi
done
