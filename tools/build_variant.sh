#!/bin/bash
# tools/build_variant.sh <name> <nvcc flags...>: builds tools/lib_<name>.so (small K,M set) with extra defines, for kbench.py
set -e
name=$1; shift
cd "$(dirname "$0")/../bayesfmmm_b200/csrc"
make -j8 OBJDIR=../../build/obj_$name OUT=../../tools/lib_$name.so EXTRA="-DBF_KM_SMALL $*" 2>&1 | grep -E "error|warning: v" || true
ls -la ../../tools/lib_$name.so
