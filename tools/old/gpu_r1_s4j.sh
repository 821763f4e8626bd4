#!/bin/bash
# session-3 GPU pass 4j (1 GPU): shuffle-scan K2 planner -- K2 tests, probe with many classes, then the full suite
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "k2 or c3 or shim" 2>&1 | tail -4 | cut -c1-300
timeout 200 python tools/probe.py k2 --rows 4000000 --dim 768 --dtype bf16 --classes 40000 2>>$O/s4j.err | cut -c1-200
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 2>>$O/s4j.err | cut -c1-200
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 200000 --zipf 2>>$O/s4j.err | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
