#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
#
# Builds the UNMODIFIED reference (marbl/Winnowmap v2.03) from the sources where they
# lie under $REF (default /root/reference) into oracle/_ref/ (git-ignored), using the
# flags of the reference's own top-level Makefile:4 / src/Makefile (we do not run the
# reference's build system: it would write objects into the read-only source tree).
#
# Products (all under oracle/_ref/):
#   winnowmap            reference + the one documented patch: `rep_len = 0` at
#                        src/map.c:281 (the reference reads rep_len uninitialised at
#                        map.c:917/933). This is THE oracle.
#   winnowmap_unpatched  the reference exactly as shipped (for reporting only).
#   libwinnowmap.a       patched objects, used by oracle/ref_harness.cpp
#   libref_harness.so    ctypes-loadable wrappers exposing ksw_extd2_sse / mm_sketch / mm_chain_dp /
#                        radix_sort_128x of the real reference to the python tests.
# No reference source is copied into the repository: the patched map.c lives only in
# oracle/_ref/build/ which is an output directory.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
B=$OUT/build
if [ ! -d "$REF/src" ]; then
  echo "[build_ref] $REF not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$B"
CXX=/usr/bin/g++   # NOT $CXX: this image exports CXX=/opt/gcc/bin/g++ which has no libgomp.spec
FLAGS="-g -Wall -O2 -DOMP_PER_READ_THREADS=1 -DHAVE_KALLOC -fopenmp -std=c++11 -Wno-sign-compare -Wno-write-strings -Wno-unused-but-set-variable -fno-tree-vectorize -w -fPIC -I$REF/src"
SRCS="kthread kalloc misc bseq sketch sdust options index chain align hit format pe esterr splitidx"
objs=""
cc() { # src obj extra...
  local src=$1 obj=$2; shift 2
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then $CXX -c $FLAGS "$@" "$src" -o "$obj"; fi
}
pids=()
for s in $SRCS; do cc "$REF/src/$s.c" "$B/$s.o" & pids+=($!); done
cc "$REF/src/ksw2_ll_sse.c" "$B/ksw2_ll_sse.o" -msse2 & pids+=($!)
for k in extz2 extd2 exts2; do
  cc "$REF/src/ksw2_${k}_sse.c" "$B/ksw2_${k}_sse41.o" -msse4.1 -DKSW_CPU_DISPATCH & pids+=($!)
  cc "$REF/src/ksw2_${k}_sse.c" "$B/ksw2_${k}_sse2.o" -msse2 -mno-sse4.1 -DKSW_CPU_DISPATCH -DKSW_SSE2_ONLY & pids+=($!)
done
cc "$REF/src/ksw2_dispatch.c" "$B/ksw2_dispatch.o" -msse4.1 -DKSW_CPU_DISPATCH & pids+=($!)
cc "$REF/src/main.c" "$B/main.o" & pids+=($!)
cc "$REF/src/map.c" "$B/map_unpatched.o" & pids+=($!)
# the one documented patch (generated file is an OUTPUT, not committed)
sed '281s/rep_len,/rep_len = 0,/' "$REF/src/map.c" > "$B/map_patched.c"
grep -q 'rep_len = 0,' "$B/map_patched.c" || { echo "[build_ref] patch did not apply" >&2; exit 1; }
cc "$B/map_patched.c" "$B/map.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
COMMON=""
for s in $SRCS ksw2_ll_sse ksw2_extz2_sse41 ksw2_extd2_sse41 ksw2_exts2_sse41 ksw2_extz2_sse2 ksw2_extd2_sse2 ksw2_exts2_sse2 ksw2_dispatch; do COMMON="$COMMON $B/$s.o"; done
rm -f "$OUT/libwinnowmap.a"; ar -csr "$OUT/libwinnowmap.a" $COMMON "$B/map.o"
$CXX -g -O2 -fopenmp "$B/main.o" $COMMON "$B/map.o" -o "$OUT/winnowmap" -lm -lz -lpthread
$CXX -g -O2 -fopenmp "$B/main.o" $COMMON "$B/map_unpatched.o" -o "$OUT/winnowmap_unpatched" -lm -lz -lpthread
if [ -f "$HERE/ref_harness.cpp" ]; then
  $CXX -O2 -fopenmp -std=c++11 -w -fPIC -shared -DHAVE_KALLOC -I"$REF/src" "$HERE/ref_harness.cpp" "$OUT/libwinnowmap.a" -o "$OUT/libref_harness.so" -lm -lz -lpthread
fi
echo "[build_ref] ok: $("$OUT/winnowmap" --version)"
