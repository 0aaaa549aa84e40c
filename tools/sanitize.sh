#!/bin/bash
# compute-sanitizer passes over a small end-to-end invocation (fixture + small synthetic batch)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 600 compute-sanitizer --tool $tool --print-limit 5 python __graft_entry__.py smoke > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|hazard|Invalid|Error" gpurun_out/sanitize_$tool.log | head -8
done
