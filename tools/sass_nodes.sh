#!/bin/bash
# how ptxas scheduled the node loops of k_walk_eval<N,0>: distance (in instructions) between successive rsqrt seeds
obj=${1:-/tmp/jxp_walk.o}; n=${2:-2}
cuobjdump -sass -fun "_ZN3jxp11k_walk_evalILi${n}ELb0EEEvNS_12WalkEvalArgsE" $obj | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/_w.sass
wc -l < /tmp/_w.sass
grep -n "MUFU.RSQ64H" /tmp/_w.sass | awk -F: 'BEGIN{p=0}{printf "%d ", $1-p; p=$1} END{print ""}'
