#!/bin/bash
# usage: run_ref.sh <datadir> <dart_ref|dart_canon> <threads> <outprefix> [extra dart flags...]
# Runs a reference binary from oracle/_ref on <datadir>/{idx,r1.fq[,r2.fq]} -> <datadir>/<outprefix>.{sam,junc,log}
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
d=$1; bin=$2; t=$3; out=$4; shift 4
if [ -f "$d/r2.fq" ]; then
  "$HERE/oracle/_ref/$bin" -i "$d/idx" -f "$d/r1.fq" -f2 "$d/r2.fq" -t "$t" -o "$d/$out.sam" -j "$d/$out.junc" "$@" > "$d/$out.log"
else
  "$HERE/oracle/_ref/$bin" -i "$d/idx" -f "$d/r1.fq" -t "$t" -o "$d/$out.sam" -j "$d/$out.junc" "$@" > "$d/$out.log"
fi
