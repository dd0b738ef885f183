#!/bin/bash
# times a profiling target (default scripts/profile_target.py) for the in-tree library and every variants/libb2pt_<v>.so
cd "$(dirname "$0")/.."
T=${TARGET:-scripts/profile_target.py}
echo "== in-tree"; python $T | tail -1
for v in "$@"; do echo "== $v"; B2PT_LIB=$PWD/variants/libb2pt_$v.so python $T | tail -1; done
