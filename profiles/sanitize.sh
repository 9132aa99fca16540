#!/bin/bash
# compute-sanitizer over profiles/san_case.py; logs under gpurun_out/san/ (copied to profiles/r02/ when clean)
mkdir -p gpurun_out/san
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/san_case.py > gpurun_out/san/$tool.log 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY' gpurun_out/san/$tool.log) summary line(s): $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/san/$tool.log | tail -1)"
done
