#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own CPU sources, from where they lie under /root/reference,
# into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot):
#   libbwaref.so   bwa_index/*.c library objects (OCC_INTV_SHIFT 7) + oracle/ref_shim.c + oracle/ref_collect_shim.c
#   libforkksw.so  src/ksw.c + oracle/fork_ksw_shim.c  (fork's ksw_extend2 with opt_ext)
#   bwa7 / bwa6    the reference's index builder compiled with OCC_INTV_SHIFT 7 / 6,
#                  i.e. the two passes of the reference's build_index.sh:46-66
#   bwa7p          bwa7 with the broken .sa reader replaced (oracle/ref_sa_shim.c): runs `bwa mem` / `bwa fastmap`
# The reference's bwt.h hard-codes OCC_INTV_SHIFT and its own build script rewrites that
# line with sed between the two passes; we do the same on a scratch copy under $TMPDIR.
# Nothing from /root/reference is copied into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/bwa_index" ]; then
  echo "[build_ref] $REF not present; keeping whatever is in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
CFLAGS="-O2 -g -fcommon -fPIC -w -DHAVE_PTHREAD"
LOBJS="utils kthread kstring ksw bwt bntseq bwa bwamem bwamem_pair bwamem_extra QSufSort bwt_gen rope rle is bwtindex"
AOBJS="bwashm bwase bwaseqio bwtgap bwtaln bamlite bwape kopen pemerge maxk bwtsw2_core bwtsw2_main bwtsw2_aux bwt_lite bwtsw2_chain fastmap bwtsw2_pair main"
for shift in 7 6; do
  D="$TMP/s$shift"; mkdir -p "$D"
  cp "$REF"/bwa_index/*.c "$REF"/bwa_index/*.h "$D"/
  sed -i "s,#define OCC_INTV_SHIFT.*,#define OCC_INTV_SHIFT $shift,g" "$D/bwt.h"
  ( cd "$D"
    for o in $LOBJS $AOBJS; do gcc -c $CFLAGS $o.c -o $o.o & done; wait
    objs=""; for o in $LOBJS $AOBJS; do objs="$objs $o.o"; done
    gcc $CFLAGS $objs -o "$OUT/bwa$shift" -lm -lz -lpthread -lrt
    if [ "$shift" = 7 ]; then
      # bwa7p: the same program with the one loader that cannot read the reference's own .sa replaced (oracle/ref_sa_shim.c):
      # the CPU `bwa mem` / `bwa fastmap` that bench.py and tools/sam_check.py time and diff against
      gcc -c $CFLAGS -Dbwt_restore_sa=bwt_restore_sa_as_shipped bwt.c -o bwt_p.o
      gcc -c $CFLAGS -I. "$HERE/ref_sa_shim.c" -o ref_sa_shim.o
      pobjs=""; for o in $LOBJS $AOBJS; do if [ "$o" = bwt ]; then pobjs="$pobjs bwt_p.o"; else pobjs="$pobjs $o.o"; fi; done
      gcc $CFLAGS $pobjs ref_sa_shim.o -o "$OUT/bwa7p" -lm -lz -lpthread -lrt
    fi
    if [ "$shift" = 7 ]; then
      # ref_collect_shim.c #includes the unmodified bwamem.c (to reach its static mem_collect_intv) and so stands in for bwamem.o
      lobjs=""; for o in $LOBJS; do [ "$o" = bwamem ] || lobjs="$lobjs $o.o"; done
      gcc -c $CFLAGS -fopenmp -I. -I"$HERE" "$HERE/ref_shim.c" -o ref_shim.o
      gcc -c $CFLAGS -I. -I"$HERE" "$HERE/ref_collect_shim.c" -o ref_collect_shim.o
      gcc -shared $CFLAGS -fopenmp $lobjs ref_shim.o ref_collect_shim.o -o "$OUT/libbwaref.so" -lm -lz -lpthread -lrt
    fi )
done
F="$TMP/fork"; mkdir -p "$F"
cp "$REF"/src/ksw.c "$REF"/src/ksw.h "$F"/
[ -f "$REF/src/malloc_wrap.h" ] && cp "$REF/src/malloc_wrap.h" "$F"/
( cd "$F"
  gcc -c $CFLAGS ksw.c -o ksw.o
  gcc -c $CFLAGS -I. "$HERE/fork_ksw_shim.c" -o shim.o
  gcc -shared $CFLAGS ksw.o shim.o -o "$OUT/libforkksw.so" -lm )
echo "[build_ref] built: $(ls "$OUT")"
# ---- fork host code (chaining, chain filter, extension-job construction): src/*.c compiled as C++ exactly
# like the reference Makefile:14 (-std=c++11 -fpermissive), on a scratch copy laid out like the reference
# tree (src/ + GASAL2/include/ = GASAL2/src/*.h, which is what GASAL2's own Makefile installs); the one
# edit is the hard-coded CUDA include path of gasal.h, the same line GASAL2/configure.sh rewrites.
M="$TMP/forkmem"; mkdir -p "$M/src" "$M/GASAL2/include"
cp "$REF"/src/*.c "$REF"/src/*.h "$M/src/"
cp -r "$REF/src/GPUSeed" "$M/src/"
cp "$REF"/GASAL2/src/*.h "$M/GASAL2/include/"
sed -i 's,#include "/usr/local/cuda[^"]*cuda_runtime.h",#include <cuda_runtime.h>,' "$M"/GASAL2/include/*.h
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
if [ -f "$CUDA_INC/cuda_runtime.h" ]; then
( cd "$M/src"
  CXXF="-O2 -g -std=c++11 -fpermissive -fPIC -w -DHAVE_PTHREAD -I$CUDA_INC -IGPUSeed"
  FOBJS="bwamem bntseq bwa utils kstring ksw bwt bwamem_pair bwamem_extra kthread malloc_wrap"
  for o in $FOBJS; do g++ -c $CXXF $o.c -o $o.o & done; wait
  g++ -c $CXXF -fopenmp -I. "$HERE/fork_mem_shim.cpp" -o fork_mem_shim.o
  objs=""; for o in $FOBJS; do objs="$objs $o.o"; done
  g++ -shared $CXXF -fopenmp $objs fork_mem_shim.o -o "$OUT/libforkmem.so" -lm -lz -lpthread )
else
  echo "[build_ref] no CUDA headers: libforkmem.so not built" >&2
fi
echo "[build_ref] built: $(ls "$OUT")"
