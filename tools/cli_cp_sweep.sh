#!/bin/bash
# developer sweep: the command line on 48 config-5a files by compression, with its phase timing
set -u
for gz in ${@:-bgzf none gzip}; do
  F2Q_CLI_TIMING=1 F2Q_DEBUG_INGEST=1 python tools/cli_ingest_bench.py --config 5a --files 48 --reads 1500000 --gz $gz > /tmp/sweep_$gz.log 2>&1
  echo "== $gz: $(tail -1 /tmp/sweep_$gz.log | cut -c1-230)"
  grep "parameters + features" /tmp/sweep_$gz.log
  grep "released" /tmp/sweep_$gz.log | sort -k8 -n | tail -3
  grep "f2q destroy" /tmp/sweep_$gz.log | sort -t, -k5 | tail -4
  grep "timing.*stream" /tmp/sweep_$gz.log | sort -k5 -n | tail -3
done
