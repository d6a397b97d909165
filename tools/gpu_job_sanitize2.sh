mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 110 compute-sanitizer --tool $tool python tools/sanitize_cases2.py 2>&1 | grep -E "sanitize cases2 done|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|Error|Traceback" | head -8
done 2>&1 | tee gpurun_out/sanitize2.txt
