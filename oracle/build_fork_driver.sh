#!/bin/bash
# oracle/build_fork_driver.sh -- TEST INFRASTRUCTURE ONLY.
# Builds the reference's own driver program (`bwa-gasal2`, all of src/*.c compiled as C++ exactly like the reference
# Makefile:14,49-50) twice, from the sources where they lie under /root/reference, on a scratch copy laid out like the
# reference tree:
#   oracle/_ref/bwa-gasal2-b200  the UNMODIFIED driver compiled against include/compat/ (installed where the driver
#                                looks for GASAL2's and GPUSeed's headers: ../GASAL2/include/ and src/GPUSeed/) and linked
#                                to bwa-mem_gpu_b200/libbwamem_b200.so instead of libgasal.a + libseed.a.  This is the
#                                drop-in of INTEGRATION.md section A, exercised for real.
#   oracle/_ref/bwa-gasal2-cpu  the same objects linked to oracle/cpu_compat.cpp instead: both boundaries on the CPU with the reference's
#                                own bwt_smem1 / bwt_sa / ksw_extend2 -- the SAM-level checker of tests/test_gpu_sam.py.
#   oracle/_ref/bwa-gasal2-ref   the same driver with the reference's own GPU libraries (GPUSeed seed_gen.cu, GASAL2
#                                *.cpp + gasal_align.cu; flags of GASAL2/run_all.sh: MAX_SEQ_LEN=153 N_CODE=4 N_PENALTY=1)
#                                compiled for sm_100 -- the reference GPU path, to diff SAM against on the GPU box.
# Only one stub is added: the SHD filter symbol (oracle/fork_driver_shim.cpp; needs Boost, never called by default).
# Nothing from /root/reference is copied into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
OUT="$HERE/_ref"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
if [ ! -d "$REF/src" ]; then echo "[build_fork_driver] $REF not present; keeping $OUT" >&2; exit 0; fi
if [ ! -f "$ROOT/bwa-mem_gpu_b200/libbwamem_b200.so" ]; then echo "[build_fork_driver] build libbwamem_b200.so first" >&2; exit 1; fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
OBJS="utils kthread kstring ksw bwt bntseq bwa bwamem bwamem_pair bwamem_extra malloc_wrap QSufSort bwt_gen rope rle is bwtindex \
 bwashm bwase bwaseqio bwtgap bwtaln bamlite bwape kopen pemerge maxk bwtsw2_core bwtsw2_main bwtsw2_aux bwt_lite bwtsw2_chain fastmap bwtsw2_pair main"
CXXF="-O3 -g -std=c++11 -fpermissive -w -msse4.2 -DHAVE_PTHREAD -DUSE_MALLOC_WRAPPERS -I$CUDA/include"

compile_host() {   # $1 = tree root holding src/ and GASAL2/include/
  ( cd "$1/src"
    for o in $OBJS; do g++ -c $CXXF -I../GASAL2/include $o.c -o $o.o & done; wait
    g++ -c -O2 "$HERE/fork_driver_shim.cpp" -o shd_stub.o
    for o in $OBJS; do [ -f $o.o ] || { echo "[build_fork_driver] $o.c failed to compile" >&2; exit 1; }; done )
}

# ---- (1) the driver over the B200 library
A="$TMP/b200"; mkdir -p "$A/src/GPUSeed" "$A/GASAL2/include"
cp "$REF"/src/*.c "$REF"/src/*.h "$A/src/"
cp "$ROOT"/include/compat/*.h "$ROOT"/include/bwamem_b200.h "$A/GASAL2/include/"
cp "$ROOT"/include/compat/seed_gen.h "$A/src/GPUSeed/seed_gen.h"
compile_host "$A"
( cd "$A/src"; objs=""; for o in $OBJS; do objs="$objs $o.o"; done
  g++ $objs shd_stub.o -o "$OUT/bwa-gasal2-b200" -L"$ROOT/bwa-mem_gpu_b200" -lbwamem_b200 \
      -Wl,-rpath,'$ORIGIN/../../bwa-mem_gpu_b200' -L"$CUDA/lib64" -lcudart -lm -lz -ldl -lpthread -lrt
  # the same objects over the CPU checker (oracle/cpu_compat.cpp: the reference's bwt_smem1 / bwt_sa from libbwaref.so and the
  # driver's own ksw_extend2): the SAM this binary writes is what the B200 one must reproduce byte for byte
  g++ -c -O2 -g -std=c++11 -w -I../GASAL2/include -IGPUSeed -I"$CUDA/include" "$HERE/cpu_compat.cpp" -o cpu_compat.o
  g++ $objs shd_stub.o cpu_compat.o -o "$OUT/bwa-gasal2-cpu" -lm -lz -ldl -lpthread -lrt )

# ---- (2) the driver over the reference's own GPU libraries, sm_100
B="$TMP/ref"; mkdir -p "$B/src" "$B/GASAL2/include" "$B/GASAL2/src"
cp "$REF"/src/*.c "$REF"/src/*.h "$B/src/"
cp -r "$REF/src/GPUSeed" "$B/src/"
cp -r "$REF"/GASAL2/src/. "$B/GASAL2/src/"
sed -i 's,#include "/usr/local/cuda[^"]*cuda_runtime.h",#include <cuda_runtime.h>,' "$B"/GASAL2/src/*.h   # what GASAL2/configure.sh rewrites
cp "$B"/GASAL2/src/*.h "$B/GASAL2/include/"
compile_host "$B"
GDEF="-DMAX_SEQ_LEN=153 -DN_CODE=4 -DN_PENALTY=1"
ARCH="-gencode arch=compute_100,code=sm_100"
( cd "$B/GASAL2/src"
  for f in args_parser host_batch ctors interfaces res; do g++ -c -g -O3 -std=c++11 -w $GDEF -I"$CUDA/include" $f.cpp -o $f.o & done
  "$CUDA/bin/nvcc" -c -O3 -std=c++11 -w -Xcompiler -w $GDEF $ARCH -lineinfo --default-stream per-thread gasal_align.cu -o gasal_align.o &
  wait )
( cd "$B/src/GPUSeed"
  "$CUDA/bin/nvcc" -c --device-c -O3 -std=c++14 -w -Xcompiler -w $ARCH -lineinfo --default-stream per-thread -I.. -I"$CUDA/include/nvtx3" seed_gen.cu -o seed_gen.o
  "$CUDA/bin/nvcc" $ARCH -dlink seed_gen.o -o dlink.o )
( cd "$B/src"; objs=""; for o in $OBJS; do objs="$objs $o.o"; done
  g++ $objs shd_stub.o GPUSeed/seed_gen.o GPUSeed/dlink.o ../GASAL2/src/*.o -o "$OUT/bwa-gasal2-ref" \
      -L"$CUDA/lib64" -lcudart -lcudadevrt -lm -lz -ldl -lpthread -lrt ) || echo "[build_fork_driver] reference GPU build failed" >&2
echo "[build_fork_driver] built: $(ls "$OUT" | tr '\n' ' ')"
