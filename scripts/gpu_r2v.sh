#!/bin/bash
# round-2 session V: compute-sanitizer on the round's new code paths (40x36x30 runs: TOPK launch + ZA tiles of the isotropic medium,
# ZA tiles + thread groups of the general anisotropic one, both free-surface routes)
OUT=gpurun_out/r2v
mkdir -p $OUT
san() { local tool=$1 name=$2 med=$3 nt=$4; shift 4; env "$@" timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_case.py $med $nt > $OUT/${tool}_$name.log 2>&1; echo "$tool $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_$name.log | tail -1) | $(grep sanitize_case $OUT/${tool}_$name.log | tail -1)"; }
san memcheck iso_fused iso 4 A=1
san memcheck iso_ktop iso 4 CGFD_FUSE_TOP=0
san memcheck aniso aniso 4 A=1
san memcheck aniso_fused aniso 4 CGFD_FUSE_TOP=1
san memcheck visco_fused visco 4 CGFD_FUSE_TOP=1
san memcheck vti_fused vti 4 CGFD_FUSE_TOP=1
san racecheck iso_fused iso 2 A=1
san racecheck aniso aniso 2 A=1
san synccheck iso_fused iso 2 A=1
san synccheck aniso_fused aniso 2 CGFD_FUSE_TOP=1
