#!/usr/bin/env python
"""CIGAR-path leg of bench.py alone (quick iteration on global_kernel): python tools/bench_cigar.py"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as ge
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--no-cpu-baseline", action="store_true")
args = ap.parse_args()
pkg = ge.load_package(); pkg.build()
torch.cuda.set_device(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
print(json.dumps(bench.run_cigar(args, pkg, flush)))
