#!/bin/bash
echo "== split (default)"; python tools/small_latency.py --reps 200 --threads 1 2>&1 | cut -c1-200
echo "== CSG_NO_SPLIT=1"; CSG_NO_SPLIT=1 python tools/small_latency.py --reps 200 --threads 1 2>&1 | cut -c1-200
