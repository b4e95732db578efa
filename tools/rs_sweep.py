"""random 32-byte-sector gather throughput vs footprint (seeding roofline denominator)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
L = pkg.lib()
for mb in (64, 100, 256, 1024, 3100, 8192):
    g = L.bwa_b200_measure_random_sector_gbs(0, mb << 20, 2048, 3)
    print(f"{mb:6d} MB  {g:8.1f} GB/s", flush=True)
