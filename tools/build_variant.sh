#!/bin/bash
# tools/build_variant.sh NAME cloud|pathtrace -DSKY_...   -> skyrendering_b200/csrc/variant_NAME.so (A/B experiments; SKYB200_LIB selects it)
set -e
cd "$(dirname "$0")/../skyrendering_b200/csrc"
name=$1; tu=$2; shift 2
nvcc -O3 -std=c++20 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -use_fast_math "$@" -Xptxas -v -dc -o /tmp/${tu}_$name.o $tu.cu 2> /tmp/${tu}_$name.log
python3 - "$name" /tmp/${tu}_$name.log <<'PY'
import re, sys
txt = open(sys.argv[2]).read()
out = []
for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*\n.*Used (\d+) registers", txt):
    n = m.group(1)
    k = re.search(r"(k19_path_traceILi3ELb0ELi1ELb0|k16_renderILi([01])ELb([01])ELb0)", n)
    if k:
        tag = "k19" if k.group(1).startswith("k19") else f"k16<M{k.group(2)},{'hw' if k.group(3) == '1' else 'exact'}>"
        out.append(f"{tag}: {m.group(4)} regs, {m.group(3)} B spill")
print(sys.argv[1], "|", "; ".join(out))
PY
objs="atmosphere.o composite.o noise.o cloud.o pathtrace.o api.o cloud_strict.o pathtrace_strict.o composite_strict.o"
objs=${objs/$tu.o//tmp/${tu}_$name.o}
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variant_$name.so $objs -cudart static
