#!/bin/bash
# A/B measurements of round 2 (measurement build: OTMB_NVCC_EXTRA="-DOTMB_AB ..."): upload pacing of the stream call
for lead in 2 3 4 100; do
  for ns in 8 16; do
    echo "== lead=$lead nslabs=$ns"
    OTMB_STREAM_LEAD=$lead timeout 120 python profiles/e2e_trace.py $ns 2>&1 | grep "ms per call"
  done
done
