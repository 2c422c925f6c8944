#!/bin/bash
# one ncu --set full capture per hot kernel of the bench configs -> gpurun_out/<tag>_<name>.ncu-rep  (run under gpurun, ONE GPU)
#   bash scripts/ncu_all.sh <tag> [name ...]     names: headline c5f64 c4pass c4topo c3 c2i8 c1 (default: all)
TAG=${1:-r02}; shift
WANT="${*:-headline c5f64 c4pass c4topo c3 c2i8 c1}"
mkdir -p gpurun_out
cap() {  # name cfg kernel-regex skip
  case " $WANT " in *" $1 "*) ;; *) return;; esac
  CFG="$2" timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$3" --launch-skip "$4" -c 1 -f -o gpurun_out/${TAG}_$1 python scripts/ncu_cfg.py > gpurun_out/${TAG}_$1.log 2>&1
  grep "attempts/launch" gpurun_out/${TAG}_$1.log
}
cap headline headline "mcg_pass_m1" 5
cap c5f64 "C5 Heisenberg sc 256^3 T-scan, fp64" "mcg_pass_m1" 5
cap c4pass "C4" "mcg_pass_m1" 5
cap c4topo "C4" "mcg_topo" 2
cap c3 "C3" "mcg_pass_m1" 9
cap c2i8 "C2 Ising square 4096^2 T-scan, int8" "mcg_pass_m1" 5
cap c1 "C1" "mcg_pass_m1" 5
