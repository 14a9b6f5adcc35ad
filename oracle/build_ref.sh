#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds the reference's OWN mapping
# code (unmodified apart from the four unguarded MAP_* macros) into
# oracle/_ref/libref_<W>x<H>_<res>.so, one shared object per grid geometry.
#
# The reference sources are never copied into git: the mapping block is cut
# out of /root/reference/uav_local_nav.c by line range at build time, written
# only under oracle/_ref/ (git-ignored; it travels to the GPU box with the
# snapshot), and compiled together with oracle/ref_shim.c.
#
#   line ranges:
#     78-82      scan-frame macros            105-110  TOF_COLS/ROWS, tof_beams_m
#     117-118    TOF_MAX_RANGE_M / TOF_FOV    182-229  grid, log-odds, world_to_grid
#     241-353    raycast, beams, recentering  356-385  frontier scoring
#     1303-1307  xor8                         1316-1359 u16 reader, robust column
#
# Flags pinned by SURVEY.md section 8(c): gcc -O2 -ffp-contract=off, no -march=native,
# no -ffast-math, dynamic glibc libm (sincosf / lrintf).
#
# usage: oracle/build_ref.sh [W:H:RES:SIZE ...]   (default: every geometry the tests use)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${UQS_REFERENCE:-/root/reference}/uav_local_nav.c"
OUT="$HERE/_ref"
WANT_SHA=d6f4a673ec58c2a54f071253e129e64ac0cf8801191fd1db4799befeee595526

if [ ! -f "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt files in $OUT" >&2
  exit 0
fi
GOT_SHA=$(sha256sum "$REF" | cut -d' ' -f1)
if [ "$GOT_SHA" != "$WANT_SHA" ]; then
  echo "build_ref: reference sha256 mismatch ($GOT_SHA) -- line ranges may be stale" >&2
  exit 1
fi
mkdir -p "$OUT"

if [ "$#" -eq 0 ]; then
  # KAT geometry (reference's native constants), C1/C3, C2, C4, and the 16 C5 resolutions
  set -- 500:500:0.10:50.0 400:400:0.05:20.0 2000:2000:0.01:20.0 16384:16384:0.01:163.84
  while read -r W RES; do set -- "$@" "$W:$W:$RES:20.0"; done < "$HERE/c5_geometries.txt"
fi

for spec in "$@"; do
  IFS=: read -r W H RES SIZE <<<"$spec"
  tag="${W}x${H}_${RES}"
  src="$OUT/ref_${tag}.c"
  lib="$OUT/libref_${tag}.so"
  if [ -f "$lib" ] && [ "$lib" -nt "$HERE/ref_shim.c" ] && [ "$lib" -nt "$0" ]; then continue; fi
  {
    printf '#include <stdio.h>\n#include <stdlib.h>\n#include <stdint.h>\n#include <stdbool.h>\n#include <string.h>\n#include <math.h>\n'
    sed -n '78,82p;105,110p;117,118p;182,229p;241,353p;356,385p;1303,1307p;1316,1359p' "$REF"
    cat "$HERE/ref_shim.c"
  } | sed -e "s/^#define MAP_RES_M .*/#define MAP_RES_M   ${RES}f/" \
          -e "s/^#define MAP_SIZE_M .*/#define MAP_SIZE_M  ${SIZE}f/" \
          -e "s/^#define MAP_W .*/#define MAP_W       ${W}/" \
          -e "s/^#define MAP_H .*/#define MAP_H       ${H}/" > "$src"
  gcc -O2 -ffp-contract=off -fPIC -shared -w -o "$lib" "$src" -lm
  rm -f "$src"   # extracted reference text never outlives the compile
  echo "built $lib"
done
