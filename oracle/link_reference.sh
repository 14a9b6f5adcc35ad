#!/usr/bin/env bash
# oracle/link_reference.sh -- TEST INFRASTRUCTURE.  Proves that include/uqs_mapping.h is a drop-in for the file it
# cites: builds the REAL /root/reference/uav_local_nav.c with its mapping block cut out and links it against
# libuqs_mapping.so (INTEGRATION.md section 2 as an executable recipe).
#
#   removed   :108 (tof_beams_m), :188-216 (grid state, log-odds constants, clamp_lo, world_to_grid, idx),
#             :229 (pending_kf_flags), :241-385 (raycast_update ... frontier_score_dir)
#   kept      :182-186 MAP_* macros, :218-227 KF_* flags, :230-239 frontier / pause constants
#   rewritten `memset(occ_grid, 0, sizeof(occ_grid));` at :2190 -> `map_reset();` (sizeof of a pointer would be 8)
#   added     #include "uqs_mapping.h" after :48; uqs_init + uqs_dropin_configure at the top of main()
#
# The un-vendored `common/mavlink.h` (:48) is replaced by oracle/stub/common/mavlink.h: types, constants and
# do-nothing inline bodies, enough for gcc -- no MAVLink behaviour is claimed.  The edited text exists only under
# oracle/_ref/link/ (git-ignored) and is deleted after the compile.  Compiled -O0 so that every call site survives into
# the object (at -O2 the reference's own `static const bool` switches let gcc drop some callers as dead code).  The program is linked, never run (it opens UARTs).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${UQS_REFERENCE:-/root/reference}/uav_local_nav.c"
OUT="$HERE/_ref/link"
WANT_SHA=d6f4a673ec58c2a54f071253e129e64ac0cf8801191fd1db4799befeee595526
if [ ! -f "$REF" ]; then
  echo "link_reference: $REF not present -- nothing to do" >&2
  exit 0
fi
GOT_SHA=$(sha256sum "$REF" | cut -d' ' -f1)
[ "$GOT_SHA" = "$WANT_SHA" ] || { echo "link_reference: reference sha256 mismatch ($GOT_SHA) -- line ranges may be stale" >&2; exit 1; }
mkdir -p "$OUT"
SRC="$OUT/uav_local_nav_cut.c"
awk '
  NR == 48  { print; print "#include \"uqs_mapping.h\"   /* B200 mapping path: occ_grid, world_to_grid, raycast_update, ... */"; next }
  NR == 108 { next }
  NR >= 188 && NR <= 216 { next }
  NR == 229 { next }
  NR >= 241 && NR <= 385 { next }
  /memset\(occ_grid, 0, sizeof\(occ_grid\)\);/ { sub(/memset\(occ_grid, 0, sizeof\(occ_grid\)\);/, "map_reset();"); print; next }
  /^int main\(int argc, char\*\* argv\) \{/ {
    print
    print "  uqs_params mp; uqs_params_default(&mp);   /* 500x500 @ 0.10 m, 4.0 m / 63 deg ToF, -1/+6/+-80 */"
    print "  mp.W = MAP_W; mp.H = MAP_H; mp.res_m = MAP_RES_M; mp.size_m = MAP_SIZE_M;"
    print "  if (uqs_init(0) || uqs_dropin_configure(&mp)) { fprintf(stderr, \"%s\\n\", uqs_last_error()); return 1; }"
    next
  }
  { print }
' "$REF" > "$SRC"
grep -q 'map_reset();' "$SRC" && grep -q 'uqs_dropin_configure' "$SRC" || { echo "link_reference: an edit did not apply" >&2; exit 1; }
gcc -O0 -std=gnu11 -c -I"$HERE/stub" -I"$ROOT/include" -o "$OUT/uav_local_nav_cut.o" "$SRC"
rm -f "$SRC"     # the edited reference text never outlives the compile
gcc -o "$OUT/uav_local_nav_b200" "$OUT/uav_local_nav_cut.o" -L"$ROOT/micro-quad-slam_b200" -luqs_mapping -lm \
    -Wl,-rpath,"$ROOT/micro-quad-slam_b200"
echo "linked $OUT/uav_local_nav_b200"
