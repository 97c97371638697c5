"""The e2e_api block of bench.py on its own (newcheb.run_variants from JSON dictionaries to gathered rows).

    python tools/e2e_api_probe.py [runs] [nt]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
for rep in range(2):
    print(json.dumps(bench.e2e_api_block(0, 1, 0, runs, nt)), flush=True)
