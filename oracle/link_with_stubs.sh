#!/bin/sh
# TEST INFRASTRUCTURE ONLY -- usage: link_with_stubs.sh <cxx> <out.so> <obj-dir>
# Links the objects into a shared library; every symbol that stays undefined (Bullet / Caffe-backed functions that nothing on the
# compiled path calls) is resolved to ref_abort_stub, so reaching one aborts the test instead of running made-up behaviour.
CXX="$1"; OUT="$2"; DIR="$3"
"$CXX" -shared -o "$OUT" "$DIR"/*.o -Wl,-z,defs -Wl,--no-demangle 2>&1 \
  | sed -n "s/.*undefined reference to \`\([^']*\)'.*/\1/p" | sort -u > "$DIR/undefined.txt"
ARGS=""
while read -r sym; do ARGS="$ARGS -Wl,--defsym,$sym=ref_abort_stub"; done < "$DIR/undefined.txt"
# shellcheck disable=SC2086
"$CXX" -shared -o "$OUT" "$DIR"/*.o $ARGS || exit 1
echo "$OUT: $(wc -l < "$DIR/undefined.txt") unused Bullet/Caffe-backed symbols -> ref_abort_stub"
