#!/bin/sh
# usage: scripts/gpusnap.sh <name> [gpurun flags...] -- freezes the working tree under gpu_snap/<name>/ (git-ignored,
# travels with the gpurun snapshot) and runs gpu_snap/<name>/scripts/job.sh there, so the tree can keep changing
# while the call waits for a GPU slot.  Results come back under gpurun_out/<name>/.
set -eu
name="$1"; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
snap="$root/gpu_snap/$name"
rm -rf "$root/gpu_snap"
mkdir -p "$snap"
( cd "$root" && tar --exclude=./.git --exclude=./gpurun_out --exclude=./gpu_snap --exclude='__pycache__' --exclude='./configurable_spectrograms_b200/build' -cf - . ) | ( cd "$snap" && tar -xf - )
exec "$root/scripts/gpurun_retry.sh" "$@" -- "cd gpu_snap/$name && mkdir -p gpurun_out && bash scripts/job.sh > gpurun_out/job.log 2>&1; rc=\$?; mkdir -p ../../gpurun_out/$name && cp -r gpurun_out/. ../../gpurun_out/$name/; tail -c 6000 gpurun_out/job.log; exit \$rc"
