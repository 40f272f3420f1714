#!/usr/bin/env bash
# round-2 GPU session 14 (1 GPU): the K6 string-name path (committed after session 13) + every indexer/trim test
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 300 -k "first_occurrence or indexer or trim or index" > gpurun_out/r2n_index_tests.log 2>&1; tail -3 gpurun_out/r2n_index_tests.log
echo done
