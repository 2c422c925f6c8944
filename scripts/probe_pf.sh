#!/bin/bash
# A/B of the bulk L2 prefetch (MCG_JIT_PF: 0 off, 1 own + neighbour rows, 2 own row only) on the configs, fp32 and fp64
for pf in ${PFS:-0 1 2}; do
  echo "== PF=$pf fp64"; MCG_JIT_PF=$pf PREC=64 ONLY="${CF64:-C5 Heisenberg,C1,C3,C4}" timeout 300 python scripts/bench_configs.py 2>&1 | grep attempts
  echo "== PF=$pf fp32"; MCG_JIT_PF=$pf ONLY="${CF32:-C5 Heisenberg,C1,C2,C3,C4}" timeout 300 python scripts/bench_configs.py 2>&1 | grep attempts
done
