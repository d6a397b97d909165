for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -E "sanitize cases done|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized" | head -8
done
