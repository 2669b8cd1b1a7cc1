#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference kernel module
# (starr/src/_sumtree_array.pyx, the Cython source of `starr._cython`) from the
# sources where they lie under /root/reference into oracle/_ref/.
#
# Recipe: cython (.pyx -> .c, into a scratch dir that is deleted afterwards) and
# then one gcc command with the flags the reference's own setup.py build uses
# (python's sysconfig CFLAGS: -O2, no -march, no -ffast-math => float32 ops are
# really float32, no FMA contraction).  We do NOT run the reference's setup.py.
#
# Output: oracle/_ref/_cython.<EXT_SUFFIX>  (git-ignored, NOT gpurun-ignored: it
# travels to the GPU box like our own .so files).  No reference source is copied
# into the repo: the generated C is written to a temp dir and removed.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${STARR_REFERENCE:-/root/reference}"
out="$here/_ref"
pyx="$ref/starr/src/_sumtree_array.pyx"
if [ ! -f "$pyx" ]; then
    echo "build_ref: $pyx not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$out"
tmp="$(mktemp -d)"
trap 'rm -rf "$tmp"' EXIT
py="${PYTHON:-python}"
suffix="$($py -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
incpy="$($py -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
incnp="$($py -c 'import numpy; print(numpy.get_include())')"
# same module name the reference's Extension("starr._cython", ...) produces
$py -m cython -3 --module-name starr._cython "$pyx" -o "$tmp/_cython.c" 2>/dev/null
gcc -shared -fPIC -fno-strict-overflow -fwrapv -DNDEBUG -O2 -w \
    -I"$incpy" -I"$incnp" "$tmp/_cython.c" -o "$out/_cython$suffix"
echo "build_ref: wrote $out/_cython$suffix"

# The experimental C++ twin (starr/experimental/src/{_sumtree_array.pyx,sumtree_array.cpp}; ext
# `starr.experimental._cython`, setup.py:56-64): two source files, cython --cplus + one g++ command.
xpyx="$ref/starr/experimental/src/_sumtree_array.pyx"
xcpp="$ref/starr/experimental/src/sumtree_array.cpp"
if [ -f "$xpyx" ] && [ -f "$xcpp" ]; then
    $py -m cython -3 --cplus --module-name starr.experimental._cython "$xpyx" -o "$tmp/_xcython.cpp" 2>/dev/null
    g++ -shared -fPIC -fno-strict-overflow -fwrapv -DNDEBUG -O2 -w \
        -I"$incpy" -I"$incnp" -I"$ref/starr/experimental/src" "$tmp/_xcython.cpp" "$xcpp" \
        -o "$out/_xcython$suffix"
    echo "build_ref: wrote $out/_xcython$suffix"
fi
