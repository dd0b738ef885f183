#!/bin/bash
# times scripts/profile_target.py for the in-tree library and every variants/libb2pt_*.so given as arguments
cd "$(dirname "$0")/.."
echo "== in-tree"; python scripts/profile_target.py | tail -1
for v in "$@"; do echo "== $v"; B2PT_LIB=$PWD/variants/libb2pt_$v.so python scripts/profile_target.py | tail -1; done
