#!/bin/bash
# Manual experiment: backward timing under the CLIPK_DBG switches (1 skip epilogue math, 2 skip TMA stores, 64 skip OUT jobs, 128 skip GRAD jobs)
for p in 0 1; do
for d in 0 1 2 3 64 128 65 129 192; do
  CLIPK_PERSISTENT=$p CLIPK_DBG=$d python tests/tools/debug_bwd_time.py 32768 2>&1 | tail -1 | sed "s/^/P=$p /"
done
done
