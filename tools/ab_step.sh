#!/bin/bash
# Same-box A/B of two builds of libadc_b200.so on the per-kernel step profile (tools/prof_step.py).
#
# The boxes of the pool differ by ~3 % in step time (clocks / power cap), more than most single kernel changes, so a change is
# judged by alternating the two libraries on ONE box:
#
#   here (no GPU):   tools/ab_step.sh build <git-ref>      # builds the library of <git-ref> into tools/_prev.so (the .so travels)
#   on the GPU box:  gpurun -- 'bash tools/ab_step.sh run ged 128 tc_glue'
#
# `run <config> <batch> <kind>` prints the `kind` rows of prof_step.py for the working-tree library and for tools/_prev.so, twice each.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/aerial_depth_completion_b200"
case "$1" in
build)
    REF="${2:-HEAD}"
    T="$(mktemp -d)"
    mkdir -p "$T/pkg/csrc" "$T/include"
    for f in $(git -C "$ROOT" ls-tree --name-only "$REF" aerial_depth_completion_b200/csrc/ | grep -E '\.(cu|cuh|h)$|Makefile$'); do
        git -C "$ROOT" show "$REF:$f" > "$T/pkg/csrc/$(basename "$f")"
    done
    git -C "$ROOT" show "$REF:include/adc_b200.h" > "$T/include/adc_b200.h"
    make -C "$T/pkg/csrc" > "$T/make.log" 2>&1 || { tail -20 "$T/make.log"; exit 1; }
    cp "$T/pkg/libadc_b200.so" "$ROOT/tools/_prev.so"
    rm -rf "$T"
    echo "tools/_prev.so = $REF"
    ;;
run)
    CFG="${2:-ged}"; B="${3:-128}"; KIND="${4:-tc_}"
    mkdir -p "$ROOT/gpurun_out"
    cp "$PKG/libadc_b200.so" /tmp/_new.so
    for rep in 1 2; do
        cp /tmp/_new.so "$PKG/libadc_b200.so"
        python "$ROOT/tools/prof_step.py" "$CFG" "$B" > "$ROOT/gpurun_out/ab_new$rep.txt" 2>&1
        echo "new:";  head -1 "$ROOT/gpurun_out/ab_new$rep.txt"; grep "$KIND" "$ROOT/gpurun_out/ab_new$rep.txt" | head -6
        cp "$ROOT/tools/_prev.so" "$PKG/libadc_b200.so"
        python "$ROOT/tools/prof_step.py" "$CFG" "$B" > "$ROOT/gpurun_out/ab_prev$rep.txt" 2>&1
        echo "prev:"; head -1 "$ROOT/gpurun_out/ab_prev$rep.txt"; grep "$KIND" "$ROOT/gpurun_out/ab_prev$rep.txt" | head -6
    done
    cp /tmp/_new.so "$PKG/libadc_b200.so"
    ;;
*)
    echo "usage: ab_step.sh build <git-ref> | run <config> <batch> <kind>"; exit 2
    ;;
esac
