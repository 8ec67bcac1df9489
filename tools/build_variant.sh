#!/bin/bash
# tools/build_variant.sh NAME cloud|pathtrace|composite|atmosphere -DSKY_...   -> skyrendering_b200/csrc/variant_NAME.so (A/B experiments; SKYB200_LIB selects it)
set -e
cd "$(dirname "$0")/../skyrendering_b200/csrc"
name=$1; tu=$2; shift 2
src=$tu.cu; extra=""
fast="-use_fast_math"
if [ "$tu" = composite ]; then src=atmosphere.cu; extra="-DSKY_COMPOSITE_TU -Xcudafe --diag_suppress=177"; fi
if [ "$tu" = atmosphere ]; then fast="-fmad=false"; fi   # the LUT translation unit: unfused fp32
nvcc -O3 -std=c++20 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr $fast $extra "$@" -Xptxas -v -dc -o /tmp/${tu}_$name.o $src 2> /tmp/${tu}_$name.log
python3 - "$name" /tmp/${tu}_$name.log <<'PY'
import re, sys
txt = open(sys.argv[2]).read()
out = []
for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*\n.*Used (\d+) registers", txt):
    n = m.group(1)
    k = re.search(r"(k19_path_traceILi3ELb0ELi1ELb0|k16_render(_coop)?ILi([01])ELb([01])E|k6_compositeILb0ELb0ELb0ELb0ELb1E)", n)
    if k:
        if k.group(1).startswith("k19"): tag = "k19"
        elif k.group(1).startswith("k6"): tag = "k6grey"
        else: tag = f"k16{'coop' if k.group(2) else ''}<M{k.group(3)},{'hw' if k.group(4) == '1' else 'exact'}>"
        out.append(f"{tag}: {m.group(4)} regs, {m.group(3)} B spill")
print(sys.argv[1], "|", "; ".join(out))
PY
objs=" atmosphere.o composite.o noise.o ibl.o earth.o cloud.o pathtrace.o api.o cloud_strict.o pathtrace_strict.o composite_strict.o"
objs=${objs/ $tu.o/ /tmp/${tu}_$name.o}
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variant_$name.so $objs -cudart static -ldl
