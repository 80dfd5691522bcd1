#!/bin/bash
# developer sweep: the command line on 48 config-5a files by compression [and --cp], with its phase timing
set -u
gz=${1:-bgzf}; shift
for cp in "$@"; do
  F2Q_CLI_TIMING=1 python tools/cli_ingest_bench.py --config 5a --files 48 --reads 1500000 --gz $gz --cp $cp > /tmp/sweep_$gz.log 2>&1
  echo "== $gz --cp $cp: $(tail -1 /tmp/sweep_$gz.log | cut -c1-200)"
  grep "parameters + features" /tmp/sweep_$gz.log
done
