#!/bin/bash
P=tools/probe/tma_probe
for x in 0 2 1 -1 63 -2; do echo "== x=$x"; $P 518 518 16 72 10 $x 0 3; done
echo "== y negative"; $P 518 518 16 72 10 0 -3 3
echo "== box 70 at 1"; $P 518 518 16 70 9 1 0 3
