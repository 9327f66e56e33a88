#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE: builds oracle/_ref/libblight_ref.so from the reference
# sources where they lie (/root/reference, read-only), through a scratch copy in $TMPDIR that carries
# the two documented one-line fixes:
#   P1 kmer.h:796,802   minimizer_naive canonicalises m-mers with k instead of m  -> canonize(mmer, m)
#   P2 kmer.h:509,511   SlidingKMer::fill pushes raw ASCII instead of 2-bit codes -> nuc2int(*it)
# Only the built .so lands in the repo tree (oracle/_ref/, git-ignored); no reference source is copied
# into the repository.  Flags follow the reference makefile:9,18 minus -flto (single TU here) and with
# -march=x86-64-v3 instead of -march=native, because the .so is built in one container and runs on
# another host (the GPU box).
set -euo pipefail
REF="${BLIGHT_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT if any" >&2
  exit 0
fi
mkdir -p "$OUT"
if [ -f "$OUT/libblight_ref.so" ] && [ "$OUT/libblight_ref.so" -nt "$HERE/ref_harness.cpp" ] \
   && [ "$OUT/libblight_ref.so" -nt "$HERE/build_ref.sh" ]; then
  exit 0
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/*.h "$REF"/*.hpp "$REF"/blight.cpp "$TMP"/
chmod u+w "$TMP"/*
sed -i '796s/canonize(mmer, k)/canonize(mmer, m)/;802s/canonize(mmer, k)/canonize(mmer, m)/' "$TMP/kmer.h"
sed -i '509s/_push_back(\*it)/_push_back(nuc2int(*it))/;511s/_push_back(\*it)/_push_back(nuc2int(*it))/' "$TMP/kmer.h"
[ "$(grep -c 'canonize(mmer, m)' "$TMP/kmer.h")" = "2" ] || { echo "build_ref: patch P1 did not apply" >&2; exit 1; }
[ "$(grep -c '_push_back(nuc2int(\*it))' "$TMP/kmer.h")" = "2" ] || { echo "build_ref: patch P2 did not apply" >&2; exit 1; }
g++ -DNDEBUG -O3 -march=x86-64-v3 -mtune=generic -std=c++11 -fopenmp -fPIC -shared -fno-access-control \
    -Wno-unused-result -I"$TMP" "$HERE/ref_harness.cpp" -o "$OUT/libblight_ref.so" -lz
echo "build_ref: built $OUT/libblight_ref.so"
