#!/bin/bash
# Builds the reference's own CLI with process_db() routed through libsqg.so:
#   * copies $REF (default /root/reference, read-only) to integration/_build/   (git-ignored)
#   * applies integration/sim_hook.patch to src/sim.c (three calls) and adds integration/sqg_host.c to src/
#   * compiles the reference's sources with its own flags and links ../squigulator_b200/libsqg.so
#   * -> oracle/_ref/squigulator_sqg  (next to the unmodified oracle/_ref/squigulator; both travel to the GPU box)
#   * copies the reference's test inputs and golden .exp files to oracle/_ref/test/ for tests/test_host_integration.py
# Nothing of the reference enters the repository's history: every output is under git-ignored directories.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
B="$HERE/_build"
rm -rf "$B"
mkdir -p "$B" "$ROOT/oracle/_ref/test"
cp -r "$REF/src" "$REF/slow5lib" "$B/"
cp "$HERE/sqg_host.c" "$B/src/"
( cd "$B" && patch -p1 < "$HERE/sim_hook.patch" )
CC="${CC:-gcc}"
CFLAGS="-g -Wall -O2 -std=c99 -I $B/slow5lib/include -I $B/src -I $ROOT/include"
OBJ=()
for f in main sim model methmodel misc thread format gensig genread ref sqg_host; do
    $CC $CFLAGS -D_GNU_SOURCE -c "$B/src/$f.c" -o "$B/$f.o"
    OBJ+=("$B/$f.o")
done
for f in slow5 slow5_idx slow5_misc slow5_press; do
    $CC -g -O2 -std=c99 -I "$B/slow5lib/include" -I "$B/slow5lib/thirdparty/streamvbyte/include" -c "$B/slow5lib/src/$f.c" -o "$B/$f.o"
    OBJ+=("$B/$f.o")
done
for f in streamvbyte_decode streamvbyte_encode streamvbyte_zigzag; do
    $CC -std=c99 -O3 -DSTREAMVBYTE_SSSE3=1 -mssse3 -I "$B/slow5lib/thirdparty/streamvbyte/include" -c "$B/slow5lib/thirdparty/streamvbyte/src/$f.c" -o "$B/$f.o"
    OBJ+=("$B/$f.o")
done
$CC "${OBJ[@]}" -o "$ROOT/oracle/_ref/squigulator_sqg" -L "$ROOT/squigulator_b200" -lsqg \
    -Wl,-rpath,'$ORIGIN/../../squigulator_b200' -lpthread -lz -rdynamic -lm
cp "$REF"/test/*.exp "$REF"/test/*.fasta "$REF"/test/*.fa "$REF"/test/*.tsv "$ROOT/oracle/_ref/test/" 2>/dev/null || true
[ -d "$REF/test/r9-models" ] && cp -r "$REF/test/r9-models" "$ROOT/oracle/_ref/test/" || true
echo "built $ROOT/oracle/_ref/squigulator_sqg"
