#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle) -- not part of the product.
#
# Builds the reference CtuCopy 4.0.2 binary from its own UNMODIFIED sources where they lie
# under /root/reference/src, linked against oracle/shim/fftw3.h (FFTW is not installed in
# this image).  Outputs go ONLY into oracle/_ref/ (git-ignored, travels with gpurun):
#     oracle/_ref/ctucopy4_O0   reference's own flags (src/subdir.mk:66: -O0)
#     oracle/_ref/ctucopy4_O2   optimised build used as the CPU baseline in bench.py
#
# One patch is unavoidable and is applied to a THROW-AWAY copy of one header in a temp dir
# (never stored in this repo): medianFilter::flush_frame() in src/vad/vad.h:156-175 falls
# off its end without a return; g++ >= 8 plants a ud2 there, so every `-vad_*` run dies
# with SIGILL.  The copy gets `return false;` appended to the else branch, which is what
# 2015-era compilers effectively did (the value is unused when ready==false,
# src/vad/vad.cc:742-745, src/io/batch.cc:243-249).
set -euo pipefail
REF=${CTU_REFERENCE_SRC:-/root/reference/src}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
# mirror the tree with symlinks; only vad/vad.h is a patched copy
mkdir -p "$TMP/src/vad"
for d in base fea io nr vdet; do ln -s "$REF/$d" "$TMP/src/$d"; done
ln -s "$REF/main.cpp" "$TMP/src/main.cpp"
ln -s "$REF/vad/vad.cc" "$TMP/src/vad/vad.cc"
python3 - "$REF/vad/vad.h" "$TMP/src/vad/vad.h" <<'EOF'
import sys, re
src = open(sys.argv[1]).read()
needle = "            ready=false;\n        }\n    }\n};"
assert src.count(needle) == 1, "vad.h layout changed; patch point not found"
open(sys.argv[2], "w").write(src.replace(needle, "            ready=false;\n            return false;\n        }\n    }\n};"))
EOF
SRCS="main.cpp io/batch.cc io/in.cc io/opts.cc io/out.cc io/pfile.cc fea/fb.cc fea/fea.cc fea/post.cc fea/fea_impl.cc fea/post_impl.cc fea/fea_trap.cc fea/fea_delta.cc nr/nr.cc vad/vad.cc"
for OPT in O0 O2; do
  mkdir -p "$TMP/obj_$OPT"
  OBJS=""
  for f in $SRCS; do
    o="$TMP/obj_$OPT/$(echo "$f" | tr '/.' '__').o"
    # vad.cc must be compiled from the temp tree so that its "vad.h" is the patched copy;
    # batch.cc / main.cpp include ../vad/vad.h relative to io/, which resolves through the
    # symlinked io/ dir to the ORIGINAL header -- medianFilter::flush_frame is only
    # odr-used from vad.cc, so that is harmless.
    g++ -$OPT -w -std=gnu++17 -I "$HERE/shim" -c "$TMP/src/$f" -o "$o" &
    OBJS="$OBJS $o"
  done
  wait
  g++ -o "$OUT/ctucopy4_$OPT" $OBJS
done
echo "build_ref: built $OUT/ctucopy4_O0 and $OUT/ctucopy4_O2"
