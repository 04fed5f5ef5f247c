#!/usr/bin/env python
"""bench.py's single-stream sections alone (batch 1, host entry points): the call-by-call stereo chain, the one-call-per-frame stereo
chain and mono extract + match, and optionally the config-4 replay.  Prints JSON.  Usage: tools/chain_bench.py [c4]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench

out = {"sequence": bench.bench_sequence(0, False)}
if "c4" in sys.argv[1:]:
    out["c4"] = bench.bench_c4(0, False)
print(json.dumps(out))
