#!/bin/bash
# How many L2 sector lookups does ONE 32-byte random load cost on this chip?  The random-sector benchmark issues a known number of
# LDG.256 (one sector each); ncu's lts__t_sectors_srcunit_tex_op_read over the same launch answers it (back_kernel shows
# 1.97 L2 sectors per algorithmic sector).  Footprints: 32 MB (fits one L2 partition), 100 MB (the C2 index), 1 GB.
set -u
mkdir -p gpurun_out
for mb in 32 100 1024; do
cat > /tmp/rs_one.py <<PY
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module("bwa-mem_gpu_b200")
print("GB/s", pkg.lib().bwa_b200_measure_random_sector_gbs(0, $mb * 1000000, 256, 1))
PY
timeout 600 ncu --clock-control none --metrics lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_sectors.sum,dram__bytes_read.sum,smsp__inst_executed_op_global_ld.sum \
   -k regex:random_sector_kernel -s 1 -c 1 --csv --log-file gpurun_out/l2_sectors_${mb}mb.csv python /tmp/rs_one.py > gpurun_out/l2_sectors_${mb}mb.log 2>&1; echo "ncu $mb MB rc=$?"
grep -E "lts__|l1tex__|dram__|smsp__" gpurun_out/l2_sectors_${mb}mb.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
