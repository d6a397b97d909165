mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 300 compute-sanitizer --tool $tool python tools/sanitize_cases3.py 2>&1 | grep -E "sanitize cases3 done|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|Error|Traceback|Assertion" | head -12
done 2>&1 | tee gpurun_out/sanitize3.txt
