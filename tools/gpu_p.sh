#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_options.py "block_order=xslab8" "block_order=xslab4,strict_loop=4" "strict_loop=2" > gpurun_out/p_check.log 2>&1; echo "check exit $?" >> gpurun_out/p_check.log
grep -E "CHECK|exit|Error" gpurun_out/p_check.log | cut -c1-300
timeout 900 python tools/ab_box.py --nb 64 --steps 30 \
  "base|strict|" "loop4|strict|strict_loop=4" "loop2|strict|strict_loop=2" "base|fast|" \
  "xs8|strict|block_order=xslab8" "xs8_loop4|strict|block_order=xslab8,strict_loop=4" "xs8_loop2|strict|block_order=xslab8,strict_loop=2" "xs8|fast|block_order=xslab8" \
  "xs4|strict|block_order=xslab4" "xs4|fast|block_order=xslab4" \
  "xs16|strict|block_order=xslab16" "xs16|fast|block_order=xslab16" > gpurun_out/p_ab.log 2>&1; echo "ab exit $?" >> gpurun_out/p_ab.log
grep -E "^AB|exit|Error" gpurun_out/p_ab.log | cut -c1-260
