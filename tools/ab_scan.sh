#!/bin/bash
# A/B timing of scan-kernel variants at full size: each line of VARIANTS is a set of environment variables for
# tools/score_once.py (the library reads them when the device handle is created).  Usage: tools/ab_scan.sh out.jsonl [args]
out=$1; shift
: > "$out"
while IFS= read -r v; do
  [ -z "$v" ] && continue
  echo "{\"variant\": \"$v\"}" >> "$out"
  env $v python tools/score_once.py "$@" >> "$out" 2>&1
done <<< "${VARIANTS:-ISSL_TRIPLE_LSUBS=1
ISSL_TRIPLE_LSUBS=2}"
