#!/bin/bash
# usage: clk.sh <op> [opt key val]
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 20 > /tmp/clk_$1.txt &
NP=$!
python tools/prof_ops.py "$@"
kill $NP
sort /tmp/clk_$1.txt | uniq -c | sort -rn | head -4
