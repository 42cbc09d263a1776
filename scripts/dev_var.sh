#!/bin/bash
# Compare development builds of libtbk (photometry_b200/lib/var/libtbk_*.so): per-kernel-class timing, one line each.
for f in photometry_b200/lib/var/libtbk_*.so; do
	echo -n "$(basename $f .so): "
	TBK_LIBPATH=$PWD/$f timeout 120 python scripts/dev_perf.py 64 2>&1 | grep -E "^chunk=64|nstreams=2 chunk=64" | tr '\n' ' '
	echo
done
